"""Byte-pair tokenizer of the OpenCLIP text tower (the ``open_clip.tokenize`` call of video_to_video/modules/embedder.py:50).

This is CLIP's "simple tokenizer" (byte-level BPE over a lower-cased, whitespace-cleaned string; 256 byte symbols + 256 end-of-word
variants + 48 894 merges + <start_of_text> / <end_of_text> = 49 408 ids; context 77, longer prompts truncated with the last id
forced to <end_of_text>).  It needs ONE data file, the merge list ``bpe_simple_vocab_16e6.txt.gz`` that ships inside the open_clip /
CLIP packages -- data, not code; pass its path.  With it (and a local ``open_clip_pytorch_model.bin``) ``FrozenOpenCLIPEmbedder`` runs
without the open_clip package.  PARITY UNPINNED: neither open_clip nor the vocabulary file is in this image; the mechanics are
unit-tested on a synthetic merge list (tests/test_text_tower.py).  ``ftfy`` is used for unicode repair when importable, as upstream does.
"""
import gzip
import html
from functools import lru_cache

import torch

try:
    import regex as _re
    _PAT = r"""<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""
except ImportError:                                           # pragma: no cover
    import re as _re
    _PAT = r"""<start_of_text>|<end_of_text>|'s|'t|'re|'ve|'m|'ll|'d|[^\W\d_]+|\d|[^\s\w]+"""


@lru_cache()
def bytes_to_unicode():
    """printable stand-ins for the 256 byte values (the bytes that are already printable map to themselves)"""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    chars, extra = keep[:], 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + extra)
            extra += 1
    return dict(zip(keep, (chr(c) for c in chars)))


def _clean(text):
    try:
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    text = html.unescape(html.unescape(text)).strip()
    return _re.sub(r"\s+", " ", text).strip()


class SimpleTokenizer:
    def __init__(self, bpe_path, context_length=77, vocab_merges=49152 - 256 - 2):
        opener = gzip.open if str(bpe_path).endswith(".gz") else open
        with opener(bpe_path, "rb") as f:
            lines = f.read().decode("utf-8").split("\n")
        merges = [tuple(m.split()) for m in lines[1:vocab_merges + 1] if len(m.split()) == 2]      # first line is a version header
        symbols = list(bytes_to_unicode().values())
        vocab = symbols + [s + "</w>" for s in symbols] + ["".join(m) for m in merges] + ["<start_of_text>", "<end_of_text>"]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.ranks = {m: i for i, m in enumerate(merges)}
        self.byte_encoder = bytes_to_unicode()
        self.cache = {"<start_of_text>": "<start_of_text>", "<end_of_text>": "<end_of_text>"}
        self.pat = _re.compile(_PAT, _re.IGNORECASE)
        self.context_length = context_length
        self.sot, self.eot = self.encoder["<start_of_text>"], self.encoder["<end_of_text>"]

    def bpe(self, token):
        """lowest-rank adjacent pair merged first, repeatedly; the last symbol of a word carries '</w>'"""
        if token in self.cache:
            return self.cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = set(zip(word[:-1], word[1:]))
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            a, b = best
            out, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == a and word[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(word[i])
                    i += 1
            word = tuple(out)
        res = " ".join(word)
        self.cache[token] = res
        return res

    def encode(self, text):
        ids = []
        for tok in self.pat.findall(_clean(text).lower()):
            tok = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(tok).split(" "))
        return ids

    def __call__(self, texts, context_length=None):
        """str | list[str] -> LongTensor (B, context_length): <start_of_text> ids <end_of_text> 0 0 ..."""
        n = context_length or self.context_length
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros(len(texts), n, dtype=torch.long)
        for i, t in enumerate(texts):
            ids = [self.sot] + self.encode(t) + [self.eot]
            if len(ids) > n:
                ids = ids[:n]
                ids[-1] = self.eot
            out[i, :len(ids)] = torch.tensor(ids)
        return out
