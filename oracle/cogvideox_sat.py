"""TEST INFRASTRUCTURE -- execute the reference's CogVideoX-5B DiT files UNMODIFIED behind a shim of SwissArmyTransformer.

STAR's DiT (cogvideox-based/) is `sat` code with two reference-owned files:
    cogvideox-based/transformer.py          the patched copy of sat/model/transformer.py (SelfAttention, MLP,
                                            BaseTransformerLayer with the LIEM gates, BaseTransformer.forward)
    cogvideox-based/sat/dit_video_concat.py DiffusionTransformer and its mixins (patch embed, 3-D rotary, adaLN layer_forward,
                                            qk-LayerNorm attention_fn, final layer)
Both are executed here exactly as they lie in the reference tree (/root/reference, or the staged git-ignored copy
oracle/_ref written by oracle/stage_reference.py).  What is NOT in the reference tree is SwissArmyTransformer==0.4.12
(cogvideox-based/sat/requirements.txt:1) itself; the pieces of it those two files import are restated below from sat's
published behaviour -- they are plumbing and textbook leaves, none of them STAR's arithmetic:
    sat.mpu.{Column,Row}ParallelLinear      -> y = x W^T + b (model-parallel size 1: sample_sr.py:263-264)
    sat.ops.layernorm.LayerNorm             -> torch.nn.LayerNorm
    sat.mpu.utils.gelu                      -> tanh-approximated GELU (unused here: the DiT passes nn.GELU('tanh'))
    sat.transformer_defaults                -> attention_forward / mlp_forward / attention_fn defaults:
                                               fused qkv split in three, (b, heads, s, 64), softmax(QK^T/sqrt d)V, dense
    sat.model.base_model.BaseModel          -> mixin registry + hook collection (`non_conflict` chaining with old_impl)
    sat.model.finetune.lora2.LoraMixin      -> W x + (alpha / r) B_i A_i x on attention.query_key_value (3 partitions) and
                                               attention.dense of every layer
    sgm.util.instantiate_from_config, sgm.modules.diffusionmodules.util (the reference's own file, loaded by path)
PINNING STATUS: the layer / model arithmetic of a17 is pinned to the reference's own files through this module; the shimmed
`sat` leaves above stay "from published behaviour" until sat can be imported (header of DESIGN.md section 4 says the same).
"""
import copy
import importlib.util
import inspect
import math
import os
import sys
import types
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ref_loader

_ns = {}


def _mod(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
        parent, _, leaf = name.rpartition(".")
        if parent:
            setattr(_mod(parent), leaf, m)
    return m


# ------------------------------------------------------------------------------------------ sat.mpu
class _Linear(nn.Module):
    def __init__(self, input_size, output_size, bias=True, params_dtype=torch.float, device=torch.device("cpu"), **unused):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(output_size, input_size, dtype=params_dtype, device=device))
        self.bias = nn.Parameter(torch.zeros(output_size, dtype=params_dtype, device=device)) if bias else None
        if device != torch.device("meta") and self.weight.device.type != "meta":
            nn.init.normal_(self.weight, std=0.02)

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class ColumnParallelLinear(_Linear):
    def __init__(self, input_size, output_size, bias=True, gather_output=True, init_method=None, stride=1,
                 keep_master_weight_for_test=False, params_dtype=torch.float, module=None, name=None, skip_init=False,
                 device=torch.device("cpu")):
        super().__init__(input_size, output_size, bias=bias, params_dtype=params_dtype, device=device)


class RowParallelLinear(_Linear):
    def __init__(self, input_size, output_size, bias=True, input_is_parallel=False, init_method=None, stride=1,
                 keep_master_weight_for_test=False, params_dtype=torch.float, module=None, name=None, skip_init=False,
                 device=torch.device("cpu"), final_bias=True):
        super().__init__(input_size, output_size, bias=bias, params_dtype=params_dtype, device=device)


class VocabParallelEmbedding(nn.Embedding):
    def __init__(self, num_embeddings, embedding_dim, params_dtype=torch.float, skip_init=False, device=torch.device("cpu")):
        super().__init__(num_embeddings, embedding_dim, dtype=params_dtype, device=device)


def _gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(0.7978845608028654 * x * (1.0 + 0.044715 * x * x)))


class LayerNorm(nn.LayerNorm):
    def __init__(self, normalized_shape, eps=1e-5, elementwise_affine=True, **unused):
        super().__init__(normalized_shape, eps=eps, elementwise_affine=elementwise_affine)


# ------------------------------------------------------------------------------------------ sat.transformer_defaults
def split_tensor_along_last_dim(tensor, num_partitions, contiguous_split_chunks=False):
    last = tensor.dim() - 1
    if isinstance(num_partitions, int):
        return torch.split(tensor, tensor.size(last) // num_partitions, dim=last)
    raise NotImplementedError("multi-query partitions are not used by the DiT")


def standard_attention(query_layer, key_layer, value_layer, attention_mask, attention_dropout=None, log_attention_weights=None,
                       scaling_attention_score=True, **kwargs):
    if scaling_attention_score:
        query_layer = query_layer / math.sqrt(query_layer.shape[-1])
    scores = torch.matmul(query_layer, key_layer.transpose(-1, -2))
    if log_attention_weights is not None:
        scores = scores + log_attention_weights
    if not (attention_mask.shape[-2] == 1 and (attention_mask > 0).all()):
        scores = torch.mul(scores, attention_mask) - 10000.0 * (1.0 - attention_mask)
    probs = F.softmax(scores, dim=-1)
    if attention_dropout is not None:
        probs = attention_dropout(probs)
    return torch.matmul(probs, value_layer)


def attention_fn_default(query_layer, key_layer, value_layer, attention_mask, attention_dropout=None, log_attention_weights=None,
                         scaling_attention_score=True, **kwargs):
    """full (mask of ones) attention: softmax(Q K^T / sqrt d) V.  Evaluated over query chunks with plain matmuls so that the
    fp32 oracle is true fp32 at N = 17 776 (a fused fp32 kernel may run TF32) and the N x N scores never materialise."""
    assert log_attention_weights is None and scaling_attention_score
    assert attention_mask.shape[-2] == 1 and bool((attention_mask > 0).all()), "the DiT runs full attention"
    b, h, n, d = query_layer.shape
    if query_layer.dtype != torch.float32:
        return F.scaled_dot_product_attention(query_layer, key_layer, value_layer)
    out = torch.empty_like(query_layer)
    kt = key_layer.transpose(-1, -2)
    rows = max(1, min(n, (2 << 30) // (4 * b * h * key_layer.shape[2])))
    for r in range(0, n, rows):
        s = torch.matmul(query_layer[:, :, r:r + rows], kt) * (d ** -0.5)
        out[:, :, r:r + rows] = torch.matmul(torch.softmax(s, dim=-1), value_layer)
    return out


def attention_forward_default(self, hidden_states, mask, **kw_args):
    self = self.transformer.layers[kw_args["layer_id"]].attention
    attention_fn = self.hooks["attention_fn"] if "attention_fn" in self.hooks else attention_fn_default
    mixed = self.query_key_value(hidden_states)
    q, k, v = split_tensor_along_last_dim(mixed, self.stride)
    dropout_fn = self.attention_dropout if self.training else None
    q, k, v = self._transpose_for_scores(q), self._transpose_for_scores(k), self._transpose_for_scores(v)
    ctx = attention_fn(q, k, v, mask, dropout_fn, **kw_args)
    ctx = ctx.permute(0, 2, 1, 3).contiguous()
    ctx = ctx.view(*ctx.size()[:-2], self.hidden_size_per_partition)
    out = self.dense(ctx)
    if self.training:
        out = self.output_dropout(out)
    return out


def mlp_forward_default(self, hidden_states, **kw_args):
    self = self.transformer.layers[kw_args["layer_id"]].mlp
    return self.dense_4h_to_h(self.activation_func(self.dense_h_to_4h(hidden_states)))


def _no_default(name):
    def fn(*a, **k):
        raise NotImplementedError(f"sat default hook {name!r} is not reached by the DiT (a mixin overrides it)")
    return fn


HOOKS_DEFAULT = {
    "attention_fn": attention_fn_default,
    "attention_forward": attention_forward_default,
    "cross_attention_forward": _no_default("cross_attention_forward"),
    "mlp_forward": mlp_forward_default,
    "word_embedding_forward": _no_default("word_embedding_forward"),
    "position_embedding_forward": _no_default("position_embedding_forward"),
    "final_forward": _no_default("final_forward"),
    "layer_forward": _no_default("layer_forward"),
}


# ------------------------------------------------------------------------------------------ sat.model
def non_conflict(func):
    func.non_conflict = True
    return func


class BaseMixin(nn.Module):
    non_conflict = non_conflict

    def __init__(self):
        super().__init__()

    def reinit(self, parent_model=None):
        pass


class BaseModel(nn.Module):
    def __init__(self, args, transformer=None, params_dtype=torch.float, **kwargs):
        super().__init__()
        self.mixins = nn.ModuleDict()
        self.collect_hooks_()
        if transformer is not None:
            self.transformer = transformer
        else:
            BaseTransformer = sys.modules["sat.model.transformer"].BaseTransformer
            self.transformer = BaseTransformer(
                num_layers=args.num_layers, vocab_size=args.vocab_size, hidden_size=args.hidden_size,
                num_attention_heads=args.num_attention_heads, max_sequence_length=args.max_sequence_length,
                embedding_dropout_prob=0.0, attention_dropout_prob=0.0, output_dropout_prob=0.0,
                inner_hidden_size=getattr(args, "inner_hidden_size", None),
                hidden_size_per_attention_head=getattr(args, "hidden_size_per_attention_head", None),
                checkpoint_activations=getattr(args, "checkpoint_activations", False),
                layernorm_epsilon=getattr(args, "layernorm_epsilon", 1e-5), layernorm_order=args.layernorm_order,
                is_decoder=getattr(args, "is_decoder", False), use_bias=getattr(args, "use_bias", True),
                use_qkv_bias=getattr(args, "use_qkv_bias", False), use_final_layernorm=getattr(args, "use_final_layernorm", True),
                hooks=self.hooks, params_dtype=params_dtype, skip_init=getattr(args, "skip_init", False),
                device=getattr(args, "device", torch.device("cpu")), **kwargs)

    def reinit(self, mixin_names=None):
        for name, m in self.mixins.items():
            if mixin_names is None or name in mixin_names:
                m.reinit(self)

    def add_mixin(self, name, new_mixin, reinit=False):
        assert name not in self.mixins
        self.mixins[name] = new_mixin
        object.__setattr__(new_mixin, "transformer", self.transformer)
        self.collect_hooks_()
        if reinit:
            new_mixin.reinit(self)

    def collect_hooks_(self):
        hooks, origins = {}, {}
        for name in HOOKS_DEFAULT:
            if hasattr(self, name):
                hooks[name], origins[name] = getattr(self, name), "model"
            for mixin_name, m in self.mixins.items():
                if not hasattr(m, name):
                    continue
                fn = getattr(m, name)
                if hasattr(fn, "non_conflict"):
                    if "old_impl" not in inspect.signature(fn).parameters:
                        raise ValueError(f"Hook {name} at {mixin_name} must accept old_impl as an argument.")
                    if name in hooks:
                        old_impl = hooks[name]
                    elif name == "attention_fn":
                        old_impl = HOOKS_DEFAULT[name]
                    else:
                        old_impl = partial(HOOKS_DEFAULT[name], self)
                    hooks[name] = partial(fn, old_impl=old_impl)
                    origins[name] = mixin_name + " -> " + origins.get(name, "default")
                elif name in hooks:
                    raise ValueError(f"Hook {name} conflicts at {mixin_name} and {origins[name]}.")
                else:
                    hooks[name], origins[name] = fn, mixin_name
        self.hooks, self.hook_origins = hooks, origins
        return hooks

    def forward(self, *args, **kwargs):
        self.transformer.hooks.clear()
        self.transformer.hooks.update(self.hooks)
        return self.transformer(*args, **kwargs)


class _LoraLinear(nn.Module):
    """sat.model.finetune.lora2.LoraLinear: original(x) + cat_i[(x A_i^T) B_i^T] * (alpha / r), one (A_i, B_i) per partition"""

    def __init__(self, original, partition, in_dim, out_dim, r, lora_alpha=1.0):
        super().__init__()
        self.original = original
        self.scaling = lora_alpha / r
        self.partition = partition
        self.matrix_A = nn.ParameterList([nn.Parameter(torch.empty(r, in_dim)) for _ in range(partition)])
        self.matrix_B = nn.ParameterList([nn.Parameter(torch.empty(out_dim // partition, r)) for _ in range(partition)])
        for p in self.matrix_A:
            nn.init.kaiming_uniform_(p, a=math.sqrt(5))
        for p in self.matrix_B:
            nn.init.zeros_(p)

    def forward(self, x):
        y = self.original(x)
        return y + torch.cat([(x @ a.T @ b.T) * self.scaling for a, b in zip(self.matrix_A, self.matrix_B)], dim=-1)


class LoraMixin(BaseMixin):
    def __init__(self, layer_num, r=0, lora_alpha=1, lora_dropout=0.0, layer_range=None, qlora=False, cross_attention=True):
        super().__init__()
        self.r, self.lora_alpha = r, lora_alpha
        self.layer_range = list(range(layer_num)) if layer_range is None else layer_range

    def reinit(self, parent_model):
        for i in self.layer_range:
            att = parent_model.transformer.layers[i].attention
            h = att.hidden_size
            att.dense = _LoraLinear(att.dense, 1, att.inner_hidden_size, h, self.r, self.lora_alpha)
            att.query_key_value = _LoraLinear(att.query_key_value, 3, h, 3 * att.inner_hidden_size, self.r, self.lora_alpha)


# ------------------------------------------------------------------------------------------ loader
def _load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def instantiate_from_config(config, **extra):
    module, cls = config["target"].rsplit(".", 1)
    return getattr(sys.modules[module] if module in sys.modules else importlib.import_module(module), cls)(
        **config.get("params", {}), **extra)


def load_reference_dit():
    """returns the reference's dit_video_concat module (with DiffusionTransformer) executed behind the sat shim"""
    if "dit" in _ns:
        return _ns["dit"]
    root = ref_loader.REF_ROOT
    tpath = os.path.join(root, "cogvideox-based", "transformer.py")
    dpath = os.path.join(root, "cogvideox-based", "sat", "dit_video_concat.py")
    if not (os.path.isfile(tpath) and os.path.isfile(dpath)):
        raise RuntimeError("CogVideoX reference files not present under %s" % root)
    mpu = _mod("sat.mpu")
    mpu.get_model_parallel_world_size = lambda: 1
    mpu.ColumnParallelLinear, mpu.RowParallelLinear, mpu.VocabParallelEmbedding = ColumnParallelLinear, RowParallelLinear, VocabParallelEmbedding
    mpu.gather_from_model_parallel_region = mpu.copy_to_model_parallel_region = lambda x: x
    mpu.checkpoint = lambda fn, *a: fn(*a)
    mu = _mod("sat.mpu.utils")
    mu.divide = lambda a, b: a // b
    mu.sqrt = math.sqrt
    mu.scaled_init_method = lambda sigma, n: (lambda t: nn.init.normal_(t, mean=0.0, std=sigma / math.sqrt(2.0 * n)))
    mu.unscaled_init_method = lambda sigma: (lambda t: nn.init.normal_(t, mean=0.0, std=sigma))
    mu.gelu = _gelu_tanh
    _mod("sat.mpu.layers").ColumnParallelLinear = ColumnParallelLinear
    ln = _mod("sat.ops.layernorm")
    ln.LayerNorm, ln.RMSNorm = LayerNorm, LayerNorm
    td = _mod("sat.transformer_defaults")
    td.HOOKS_DEFAULT, td.standard_attention, td.attention_fn_default = HOOKS_DEFAULT, standard_attention, attention_fn_default
    td.split_tensor_along_last_dim = split_tensor_along_last_dim
    bm = _mod("sat.model.base_model")
    bm.BaseModel, bm.non_conflict = BaseModel, non_conflict
    _mod("sat.model.mixins").BaseMixin = BaseMixin
    _mod("sat.model.finetune.lora2").LoraMixin = LoraMixin
    _mod("sgm.util").instantiate_from_config = instantiate_from_config

    class Timestep(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
    _mod("sgm.modules.diffusionmodules.openaimodel").Timestep = Timestep
    _load_by_path("sgm.modules.diffusionmodules.util",
                  os.path.join(root, "cogvideox-based", "sat", "sgm", "modules", "diffusionmodules", "util.py"))
    _load_by_path("sat.model.transformer", tpath)                 # STAR's patched copy *is* sat/model/transformer.py
    dit = _load_by_path("dit_video_concat", dpath)
    _ns["dit"] = dit
    return dit


def build_reference_dit(num_layers=42, hidden_size=3072, num_attention_heads=48, num_frames=49, latent_height=60, latent_width=90,
                        text_length=226, text_hidden_size=4096, lora_r=512, time_embed_dim=512, in_channels=16, out_channels=16,
                        device="cpu"):
    """DiffusionTransformer exactly as configs/cogvideox_5b/cogvideox_5b_infer_sr.yaml builds it (network_config), with
    overridable sizes for reduced tests.  Parameters are left at the modules' own initialisation; load a state dict next."""
    dit = load_reference_dit()
    targs = types.SimpleNamespace(checkpoint_activations=False, vocab_size=1, max_sequence_length=64, layernorm_order="pre",
                                  skip_init=False, model_parallel_size=1, is_decoder=False, device=torch.device(device))
    modules = {
        "pos_embed_config": {"target": "dit_video_concat.Rotary3DPositionEmbeddingMixin",
                             "params": {"hidden_size_head": hidden_size // num_attention_heads, "text_length": text_length}},
        "patch_embed_config": {"target": "dit_video_concat.ImagePatchEmbeddingMixin", "params": {"text_hidden_size": text_hidden_size}},
        "adaln_layer_config": {"target": "dit_video_concat.AdaLNMixin", "params": {"qk_ln": True}},
        "final_layer_config": {"target": "dit_video_concat.FinalLayerMixin"},
    }
    if lora_r:
        modules["lora_config"] = {"target": "sat.model.finetune.lora2.LoraMixin", "params": {"r": lora_r}}
    net = dit.DiffusionTransformer(
        transformer_args=targs, num_frames=num_frames, time_compressed_rate=4, latent_width=latent_width, latent_height=latent_height,
        patch_size=2, in_channels=in_channels, out_channels=out_channels, hidden_size=hidden_size, num_layers=num_layers,
        num_attention_heads=num_attention_heads, elementwise_affine=True, time_embed_dim=time_embed_dim, adm_in_channels=256,
        modules=copy.deepcopy(modules))
    return net.eval().requires_grad_(False)
