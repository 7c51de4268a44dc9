"""Pipeline entry of STAR video super-resolution on star_b200.

Mirrors ``VideoToVideo_sr`` of the reference's video_to_video/video_to_video_model.py
(:20-139: constructor, ``test``; :141-161 VAE helpers; :164-210 ``pad_to_fit`` /
``make_chunks`` / ``sliding_windows_1d``).  What differs:

* the denoiser is star_b200's ``ControlledV2VUNet`` (sm_100a kernels), not autocast PyTorch;
* the diffusion block of ``test`` (ref :98-123) is factored into ``denoise_latents`` so that
  the latent-in / latent-out hot path can be driven (and benchmarked) without the VAE;
* the temporal VAE is star_b200's own ``AutoencoderKLTemporalDecoder`` (modules/temporal_vae.py, same
  checkpoint layout as diffusers', sm_100a kernels), loaded from ``opt.vae_path`` (no hub download);
* the text encoder (open_clip) is an un-vendored third-party package: it is imported lazily, and
  ready-made objects can be injected (``text_encoder=``, ``vae=``, ``generator=``), which is how the
  tests and the benchmark run on boxes without those packages or checkpoints;
* with ``torch.distributed`` initialised (one process per GPU) the whole entry is sharded:
  - each rank uploads, upsamples and VAE-encodes only its contiguous share of the frames; the latents (0.42 MB per
    frame) are all-gathered, so every rank conditions on the SAME posterior sample (the reference draws it once);
  - frame chunks -- or (chunk, CFG branch) pairs when there are at least two ranks per chunk -- are sharded across
    ranks with one all-gather per solver step (diffusion_sdedit.GaussianDiffusion.sample_sr, ``chunk_parallel``);
  - each rank decodes its share of the 3-frame VAE windows and ONE all-gather of the decoded, cropped frames
    assembles the clip (the north-star's end collective).
"""
from typing import Any, Dict

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .. import ops
from .diffusion.diffusion_sdedit import GaussianDiffusion
from .diffusion.schedules_sdedit import noise_schedule
from .modules.unet_v2v import ControlledV2VUNet, rearrange
from .utils.config import cfg
from .utils.logger import get_logger

logger = get_logger()

__all__ = ["VideoToVideo_sr", "pad_to_fit", "make_chunks", "sliding_windows_1d"]


def _dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def _shard_bounds(n, world):
    """contiguous, balanced split of n items over `world` ranks: [(lo, hi)] (empty shards when n < world)"""
    base, extra = divmod(n, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _all_gather_varlen(local, counts, dim):
    """all-gather of per-rank tensors whose extent along `dim` is counts[rank] (padded to the longest share);
    returns the concatenation in rank order.  ONE collective."""
    world, rank = _dist_info()
    pad = max(counts)
    shape = list(local.shape)
    shape[dim] = pad
    send = local.new_zeros(shape)
    if counts[rank]:
        send.narrow(dim, 0, counts[rank]).copy_(local)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send)
    return torch.cat([recv[r].narrow(dim, 0, counts[r]) for r in range(world) if counts[r]], dim=dim)


class VideoToVideo_sr():
    def __init__(self, opt, device=torch.device('cuda:0'), text_encoder=None, vae=None, generator=None, cuda_graph=False):
        self.opt = opt
        self.device = device
        self.cuda_graph = cuda_graph           # serve the denoiser's CFG-pair call from CUDA graphs (launch-bound small latents)

        if text_encoder is None:
            from .modules.embedder import FrozenOpenCLIPEmbedder
            # local files instead of the open_clip package when the options name them (weights: open_clip_pytorch_model.bin of
            # laion/CLIP-ViT-H-14-laion2B-s32B-b79K, merges: open_clip's bpe_simple_vocab_16e6.txt.gz)
            text_encoder = FrozenOpenCLIPEmbedder(device=self.device, pretrained="laion2b_s32b_b79k",
                                                  weights_path=getattr(opt, "clip_weights_path", None),
                                                  bpe_path=getattr(opt, "clip_bpe_path", None))
            text_encoder.model.to(self.device)
            logger.info('Build encoder with FrozenOpenCLIPEmbedder')
        self.text_encoder = text_encoder

        if generator is None:
            generator = ControlledV2VUNet().to(self.device).eval()
            cfg.model_path = opt.model_path
            load_dict = torch.load(cfg.model_path, map_location='cpu')
            if 'state_dict' in load_dict:
                load_dict = load_dict['state_dict']
            ret = generator.load_state_dict(load_dict, strict=False)
            logger.info('Load model path {}, with local status {}'.format(cfg.model_path, ret))
        self.generator = generator.half()
        self._graphed = None

        sigmas = noise_schedule(schedule='logsnr_cosine_interp', n=1000, zero_terminal_snr=True,
                                scale_min=2.0, scale_max=4.0)
        self.diffusion = GaussianDiffusion(sigmas=sigmas)

        if vae is None:
            # ref :57-63 downloads stabilityai/stable-video-diffusion-img2vid (subfolder vae, variant fp16) through
            # diffusers; here the same checkpoint is read from a local snapshot into the sm_100a temporal VAE.
            from .modules.temporal_vae import AutoencoderKLTemporalDecoder
            vae_path = _find_vae_dir(opt)
            if vae_path is None:
                raise ValueError("VideoToVideo_sr: no temporal-VAE weights found -- pass opt.vae_path, set STAR_VAE_PATH or "
                                 "keep the snapshot the reference downloads (stabilityai/stable-video-diffusion-img2vid, "
                                 "subfolder vae) in the Hugging Face cache; or pass vae=...  (denoise_latents() needs neither)")
            vae = AutoencoderKLTemporalDecoder.from_pretrained(vae_path, variant="fp16")
            vae.eval()
            vae.requires_grad_(False)
            vae.to(self.device)
        self.vae = vae

        self.negative_prompt = cfg.negative_prompt
        self.positive_prompt = cfg.positive_prompt
        self.negative_y = self._encode_text(self.negative_prompt)

    def _encode_text(self, y):
        """str -> (1,77,1024) via the text encoder; tensors pass through (precomputed embedding)."""
        if torch.is_tensor(y):
            return y.to(self.device)
        return self.text_encoder(y).detach()

    # -- latent-space hot path (ref :98-123) ------------------------------------------------------
    @torch.no_grad()
    def denoise_latents(self, video_data_feature, y, negative_y=None, total_noise_levels=1000, steps=50,
                        solver_mode='fast', guide_scale=7.5, max_chunk_len=32, noise=None, noise_sampler=None,
                        chunk_parallel="auto"):
        """video_data_feature: (1,4,F,h,w) VAE latent of the upsampled LR clip (any device; moved to
        self.device), y / negative_y: (1,77,1024).  Returns the denoised latent (1,4,F,h,w) fp32 on
        self.device.  ``noise`` / ``noise_sampler`` pin the two random inputs (diffuse noise, SDE noise)."""
        feat = self._upload_frames(video_data_feature)
        y = y.to(self.device)
        negative_y = (self.negative_y if negative_y is None else negative_y).to(self.device)
        frames_num = feat.shape[2]
        t = torch.LongTensor([total_noise_levels - 1]).to(self.device)
        noised_lr = self.diffusion.diffuse(feat, t, noise=noise).contiguous()
        if noise is None and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1:
            torch.distributed.broadcast(noised_lr, 0)       # chunk-parallel ranks start from the same sample
        model_kwargs = [{'y': y}, {'y': negative_y}, {'hint': feat}]
        chunk_inds = make_chunks(frames_num, interp_f_num=0, max_chunk_len=max_chunk_len) \
            if frames_num > max_chunk_len else None
        extra = {} if noise_sampler is None else {'noise_sampler': noise_sampler}
        model = self.generator
        if self.cuda_graph and hasattr(model, "forward_cfg_pair"):
            if self._graphed is None:
                from .cuda_graph import GraphedCFGPair
                self._graphed = GraphedCFGPair(self.generator)
            model = self._graphed
        return self.diffusion.sample_sr(
            noise=noised_lr, model=model, model_kwargs=model_kwargs, guide_scale=guide_scale,
            guide_rescale=0.2, solver='dpmpp_2m_sde', solver_mode=solver_mode, return_intermediate=None,
            steps=steps, t_max=total_noise_levels - 1, t_min=0, discretization='trailing',
            chunk_inds=chunk_inds, chunk_parallel=chunk_parallel, **extra)

    def _upload_frames(self, x, dim=2):
        """Host tensor -> fp32 on self.device.  Under torch.distributed every rank uploads only its share of the
        frames over its own PCIe link and the shares are all-gathered over NVLink (instead of W full uploads)."""
        world, rank = _dist_info()
        if x.is_cuda or world == 1 or x.shape[dim] < world:
            return x.to(self.device, torch.float32)
        bounds = _shard_bounds(x.shape[dim], world)
        lo, hi = bounds[rank]
        local = x.narrow(dim, lo, hi - lo).contiguous().to(self.device, torch.float32)
        return _all_gather_varlen(local, [b - a for a, b in bounds], dim)

    # -- pixel-space entry (ref :75-139) ------------------------------------------------------------
    @torch.no_grad()
    def test(self, input: Dict[str, Any], total_noise_levels=1000, steps=50, solver_mode='fast', guide_scale=7.5,
             max_chunk_len=32):
        """The reference entry (ref :75-139): (1, 3, F, H, W) fp32 on the CPU."""
        return self._test_on_device(input, total_noise_levels, steps, solver_mode, guide_scale, max_chunk_len).type(torch.float32).cpu()

    @torch.no_grad()
    def enhance_frames(self, input: Dict[str, Any], total_noise_levels=1000, steps=50, solver_mode='fast', guide_scale=7.5,
                       max_chunk_len=32, uint8=True):
        """test() + the CLI's post-processing (inference_sr.py:77-80: tensor2vid, adain_color_fix against the LR clip) with the
        post-processing on the GPU: returns (F, H, W, 3) frames in [0, 255] on the CPU -- uint8 by default, i.e. a quarter of the
        device->host traffic of test()'s fp32 tensor."""
        vid = self._test_on_device(input, total_noise_levels, steps, solver_mode, guide_scale, max_chunk_len).float()
        if vid.is_cuda:
            return ops.adain_color_fix(vid, input['video_data'].to(vid.device).float(), uint8=uint8).cpu()
        raise NotImplementedError("enhance_frames runs the GPU post-processing kernels; use test() + the reference's CPU helpers otherwise")

    def _test_on_device(self, input, total_noise_levels, steps, solver_mode, guide_scale, max_chunk_len):
        video_data = input['video_data']
        y = input['y']
        (target_h, target_w) = input['target_res']
        world, rank = _dist_info()
        frames_num = video_data.shape[0]
        bounds = _shard_bounds(frames_num, world)
        lo, hi = bounds[rank]
        # this rank's frames only: upload (LR frames), bilinear x4 + pad-with-1 on the GPU, encode frame by frame
        padding = pad_to_fit(target_h, target_w)
        h, w = target_h, target_w
        local = video_data[lo:hi].to(self.device)
        logger.info(f'video_data shape: {(frames_num, video_data.shape[1], h, w)}')
        bs = 1
        if hi > lo:
            local = self.upsample_pad(local, target_h, target_w, padding)
            feat_local = self.vae_encode(local.unsqueeze(0))
        else:
            lh = (h + padding[2] + padding[3]) // 8
            lw = (w + padding[0] + padding[1]) // 8
            feat_local = torch.zeros((1, 4, 0, lh, lw), device=self.device)
        del local
        video_data_feature = feat_local if world == 1 else \
            _all_gather_varlen(feat_local.float(), [b - a for a, b in bounds], 2)
        y = self._encode_text(y)
        gen_vid = self.denoise_latents(video_data_feature, y, None, total_noise_levels, steps, solver_mode,
                                       guide_scale, max_chunk_len)
        logger.info('sampling, finished.')
        w1, w2, h1, h2 = padding
        with torch.autocast('cuda', enabled=torch.cuda.is_available()):
            if world == 1:
                vid_tensor_gen = self.vae_decode_chunk(gen_vid, chunk_size=3)[:, :, h1:h + h1, w1:w + w1]
            else:
                vid_tensor_gen = self._decode_sharded(gen_vid, 3, (h1, h + h1, w1, w + w1))
        logger.info('temporal vae decoding, finished.')
        return rearrange(vid_tensor_gen, '(b f) c h w -> b c f h w', b=bs)

    def upsample_pad(self, frames, target_h, target_w, padding):
        """(f,3,h,w) -> bilinear resize to (target_h, target_w) (F.interpolate semantics, align_corners=False) and
        constant-1 padding (w1, w2, h1, h2), ref :81-87 -- one kernel on the GPU, eager torch for CPU tensors."""
        if frames.is_cuda:
            return ops.bilinear_pad(frames.float(), target_h, target_w, padding, 1.0)
        return F.pad(F.interpolate(frames, [target_h, target_w], mode='bilinear'), padding, 'constant', 1)

    def _decode_sharded(self, z, chunk_size, crop):
        """3-frame VAE windows (ref :144-151) round-robin over the ranks; ONE all-gather of the decoded, cropped
        frames.  Window w covers frames [3w, 3w+3) of the full clip on every rank, so the windows -- and therefore
        the decoder's per-window GroupNorm / temporal-conv context -- are exactly the single-GPU ones."""
        world, rank = _dist_info()
        t0, t1, l0, l1 = crop
        zf = rearrange(z, "b c f h w -> (b f) c h w")
        n = zf.shape[0]
        starts = list(range(0, n, chunk_size))
        mine = starts[rank::world]
        outs = [self.temporal_vae_decode(zf[s:s + chunk_size], min(chunk_size, n - s))[:, :, t0:t1, l0:l1] for s in mine]
        counts = [sum(min(chunk_size, n - s) for s in starts[r::world]) for r in range(world)]
        if outs:
            local = torch.cat(outs)
        else:
            local = torch.zeros((0, 3, t1 - t0, l1 - l0), device=self.device, dtype=torch.float16)
        dtype_code = torch.tensor([0 if local.dtype == torch.float16 else 1], device=self.device)
        dist.all_reduce(dtype_code, op=dist.ReduceOp.MAX)
        local = local.to(torch.float16 if int(dtype_code.item()) == 0 else torch.float32)
        gathered = _all_gather_varlen(local, counts, 0)                   # rank-major: windows r, r+W, ... of rank r
        order, pos = [], 0
        frame_of = {}
        for r in range(world):
            for s in starts[r::world]:
                k = min(chunk_size, n - s)
                for j in range(k):
                    frame_of[s + j] = pos + j
                pos += k
        idx = torch.tensor([frame_of[f] for f in range(n)], device=gathered.device)
        return gathered.index_select(0, idx)

    # -- VAE helpers (ref :141-161): exact 3-frame decode windows, 1-frame encode ----------------------
    def temporal_vae_decode(self, z, num_f):
        return self.vae.decode(z / self.vae.config.scaling_factor, num_frames=num_f).sample

    def vae_decode_chunk(self, z, chunk_size=3):
        z = rearrange(z, "b c f h w -> (b f) c h w")
        video = []
        for ind in range(0, z.shape[0], chunk_size):
            num_f = z[ind:ind + chunk_size].shape[0]
            video.append(self.temporal_vae_decode(z[ind:ind + chunk_size], num_f))
        return torch.cat(video)

    def vae_encode(self, t, chunk_size=1):
        num_f = t.shape[1]
        t = rearrange(t, "b f c h w -> (b f) c h w")
        z_list = []
        for ind in range(0, t.shape[0], chunk_size):
            z_list.append(self.vae.encode(t[ind:ind + chunk_size]).latent_dist.sample())
        z = rearrange(torch.cat(z_list, dim=0), "(b f) c h w -> b c f h w", f=num_f)
        return z * self.vae.config.scaling_factor


def _find_vae_dir(opt):
    """opt.vae_path / cfg.vae_path / $STAR_VAE_PATH, else the snapshot the reference's from_pretrained call (ref :57-59)
    leaves in the Hugging Face cache -- so that the reference's unmodified CLI, which only sets model_path, keeps working."""
    import glob
    import os
    for cand in (getattr(opt, 'vae_path', None), getattr(cfg, 'vae_path', None), os.environ.get('STAR_VAE_PATH')):
        if cand:
            return cand
    hub = os.environ.get('HF_HUB_CACHE') or os.path.join(os.environ.get('HF_HOME', os.path.expanduser('~/.cache/huggingface')), 'hub')
    hits = sorted(glob.glob(os.path.join(hub, 'models--stabilityai--stable-video-diffusion-img2vid', 'snapshots', '*', 'vae')))
    return hits[-1] if hits else None


def _centre_pad(size, target):
    lo = int((target - size) // 2)
    return lo, target - lo - size


def pad_to_fit(h, w):
    """(w1, w2, h1, h2) for F.pad (ref :164-187): small inputs are centred in 720x1280; larger ones
    are padded bottom/right so that H = 16 (mod 64) and W = 0 (mod 64) -> latent H = 2 (mod 8)."""
    best_h, best_w = 720, 1280
    if h < best_h:
        h1, h2 = _centre_pad(h, best_h)
    elif h == best_h:
        h1 = h2 = 0
    else:
        h1, h2 = 0, int((h + 48) // 64 * 64) + 64 - 48 - h
    if w < best_w:
        w1, w2 = _centre_pad(w, best_w)
    elif w == best_w:
        w1 = w2 = 0
    else:
        w1, w2 = 0, int(w // 64 * 64) + 64 - w
    return (w1, w2, h1, h2)


def sliding_windows_1d(length, window_size, overlap_size):
    """Windows of ``window_size`` with stride window-overlap; the last window absorbs the remainder
    when fewer than 1.25 windows are left (ref :199-210)."""
    stride = window_size - overlap_size
    coords, ind = [], 0
    while ind < length:
        if ind + window_size * 1.25 >= length:
            coords.append((ind, length))
            break
        coords.append((ind, ind + window_size))
        ind += stride
    return coords


def make_chunks(f_num, interp_f_num, max_chunk_len, chunk_overlap_ratio=0.5):
    """Chunk index list for a clip of f_num frames (ref :190-196)."""
    max_o_len = max_chunk_len * chunk_overlap_ratio
    chunk_len = int((max_chunk_len - 1) // (1 + interp_f_num) * (interp_f_num + 1) + 1)
    o_len = int((max_o_len - 1) // (1 + interp_f_num) * (interp_f_num + 1) + 1)
    return sliding_windows_1d(f_num, chunk_len, o_len)
