"""CogVideoX 3-D causal VAE (cogvideox-based/sat/vae_modules/cp_enc_dec.py: ``ContextParallelDecoder3D`` :839-983 and, at the end
of this file, ``ContextParallelEncoder3D`` :716-836) on the star_b200 kernels -- SURVEY section 8 row f4: the decode tail of the
CogVideoX path (sample_sr.py:206-230) and the encode of the LQ clip in front of it (diffusion_video.py:279-282).

Parameter tree = the reference decoder's (``first_stage_model.decoder.*`` of the 3d-vae checkpoint):
    conv_in.conv.*, conv_out.conv.*                                   ContextParallelCausalConv3d 3x3x3            (:360-430)
    mid.block_{1,2}.*, up.{l}.block.{j}.*                             ContextParallelResnetBlock3D                 (:614-713)
        .norm{1,2}.{norm_layer,conv_y.conv,conv_b.conv}.*             SpatialNorm3D (GroupNorm32 * conv_y(zq) + conv_b(zq)) (:451-510)
        .conv{1,2}.conv.*, .nin_shortcut.*
    up.{l}.upsample.conv.*                                            Upsample3D (nearest x2 [+ time x2], Conv2d)  (:531-568)
    norm_out.*                                                        SpatialNorm3D
``nn`` only stores the parameters.  Execution: every activation is a token matrix X[(t h w), C] (channels-last, one clip).

* causal Conv3d 3x3x3 = ONE 27-tap implicit GEMM (``star_conv3d_causal``): the conv's input is written by its producer
  (SpatialNorm + SiLU) straight into a buffer with two leading frames -- the temporal context the reference concatenates in
  front (copies of frame 0 for the first latent chunk, the cached last two input frames of the previous chunk afterwards,
  cp_enc_dec.py:265-268,:401-425).  The cache stays ON THE GPU: the reference moves it to the CPU and back for every conv of
  every chunk (``.clone().cpu()`` :408-410, ``.to(input_.device)`` :266).
* SpatialNorm3D: conv_y / conv_b are 1x1x1 convolutions of zq, so they commute with zq's nearest-neighbour interpolation: both
  are ONE small GEMM per norm at latent resolution ([Tl*Hl*Wl, 2C]) and ``star_groupnorm_mod`` gathers them per row while it
  applies the GroupNorm affine and the SiLU that always follows -- the reference materialises two full-resolution tensors.
* residual adds ride in the second conv's epilogue; the 1x1x1 shortcut is a GEMM.
Decoding follows the reference's chunk protocol exactly (``decode(z_chunk, clear_fake_cp_cache=...)``; GroupNorm statistics are
per chunk there, so they are here): ``decode_latent`` mirrors the loop of sample_sr.py:212-227.
"""
import torch
import torch.nn as nn

from .. import ops


class _CausalConv3d(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, k)


class _SpatialNorm3D(nn.Module):
    def __init__(self, c, zq_ch):
        super().__init__()
        self.norm_layer = nn.GroupNorm(32, c, eps=1e-6, affine=True)
        self.conv_y = _CausalConv3d(zq_ch, c, 1)
        self.conv_b = _CausalConv3d(zq_ch, c, 1)


class _ResnetBlock3D(nn.Module):
    def __init__(self, cin, cout, zq_ch):
        super().__init__()
        self.norm1 = _SpatialNorm3D(cin, zq_ch)
        self.conv1 = _CausalConv3d(cin, cout, 3)
        self.norm2 = _SpatialNorm3D(cout, zq_ch)
        self.conv2 = _CausalConv3d(cout, cout, 3)
        if cin != cout:
            self.nin_shortcut = nn.Conv3d(cin, cout, 1)


class _ResnetBlock3DPlain(nn.Module):
    """encoder block: plain GroupNorm32 (``Normalize``, cp_enc_dec.py:444-448) instead of SpatialNorm3D"""

    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6, affine=True)
        self.conv1 = _CausalConv3d(cin, cout, 3)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6, affine=True)
        self.conv2 = _CausalConv3d(cout, cout, 3)
        if cin != cout:
            self.nin_shortcut = nn.Conv3d(cin, cout, 1)


class _DownSample3D(nn.Module):
    def __init__(self, c, compress_time):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)
        self.compress_time = compress_time


class _Upsample3D(nn.Module):
    def __init__(self, c, compress_time):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)
        self.compress_time = compress_time


ZQ_PAD = 64          # latent channels are zero-padded to one 64-wide K chunk of the GEMM kernels


class ContextParallelDecoder3D(nn.Module):
    """Same constructor keywords, state-dict keys and ``forward(z, clear_fake_cp_cache=True)`` as the reference class
    (context-parallel size 1, ``add_conv=False``, no attention levels: the shipped configuration, yaml :128-141)."""

    def __init__(self, *, ch=128, out_ch=3, ch_mult=(1, 2, 2, 4), num_res_blocks=3, attn_resolutions=(), dropout=0.0,
                 resamp_with_conv=True, in_channels=3, resolution=256, z_channels=16, give_pre_end=False, zq_ch=None,
                 add_conv=False, pad_mode="first", temporal_compress_times=4, gather_norm=False, **ignorekwargs):
        super().__init__()
        if add_conv or give_pre_end or len(attn_resolutions) or not resamp_with_conv:
            raise NotImplementedError("only the shipped CogVideoX decoder configuration is built")
        assert z_channels <= ZQ_PAD and out_ch <= 8
        self.ch, self.out_ch, self.z_channels = ch, out_ch, z_channels
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        tlevel = {1: 0, 2: 1, 4: 2, 8: 3}[temporal_compress_times]
        zq_ch = z_channels if zq_ch is None else zq_ch
        block_in = ch * ch_mult[-1]
        self.conv_in = _CausalConv3d(z_channels, block_in, 3)
        self.mid = nn.Module()
        self.mid.block_1 = _ResnetBlock3D(block_in, block_in, zq_ch)
        self.mid.block_2 = _ResnetBlock3D(block_in, block_in, zq_ch)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            up = nn.Module()
            up.block, up.attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                up.block.append(_ResnetBlock3D(block_in, block_out, zq_ch))
                block_in = block_out
            if i_level != 0:
                up.upsample = _Upsample3D(block_in, compress_time=not (i_level < self.num_resolutions - tlevel))
            self.up.insert(0, up)
        self.norm_out = _SpatialNorm3D(block_in, zq_ch)
        self.conv_out = _CausalConv3d(block_in, out_ch, 3)
        self._packed = None
        self._cache = {}                 # causal-conv context frames carried from one latent chunk to the next (device tensors)

    # ---- packing -------------------------------------------------------------------------------------------------------
    def _dtype(self):
        dt = self.conv_in.conv.weight.dtype
        return dt if dt in (torch.float16, torch.bfloat16) else torch.float16

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _pack(self):
        dt, dev = self._dtype(), self.conv_in.conv.weight.device

        def h(t):
            return t.detach().to(device=dev, dtype=dt).contiguous()

        def w27(c, pad_in=None):                                   # [Cout, Cin, t, h, w] -> [Cout, t, h, w, Cin(pad)]
            w = c.conv.weight.detach().float().permute(0, 2, 3, 4, 1)
            if pad_in:
                w = torch.nn.functional.pad(w, (0, pad_in - w.shape[-1]))
            return h(w), h(c.conv.bias)

        def norm(n):                                               # conv_y | conv_b as one [2C, ZQ_PAD] GEMM weight
            wy, wb = n.conv_y.conv.weight.detach().float()[:, :, 0, 0, 0], n.conv_b.conv.weight.detach().float()[:, :, 0, 0, 0]
            w = torch.nn.functional.pad(torch.cat([wy, wb], 0), (0, ZQ_PAD - wy.shape[1]))
            b = torch.cat([n.conv_y.conv.bias.detach().float(), n.conv_b.conv.bias.detach().float()], 0)
            return {"g": h(n.norm_layer.weight), "b": h(n.norm_layer.bias), "w": h(w), "wb": h(b), "C": wy.shape[0]}

        def res(r):
            p = {"n1": norm(r.norm1), "c1": w27(r.conv1), "n2": norm(r.norm2), "c2": w27(r.conv2)}
            if hasattr(r, "nin_shortcut"):
                p["nin"] = (h(r.nin_shortcut.weight.detach()[:, :, 0, 0, 0]), h(r.nin_shortcut.bias))
            return p

        pk = {"in": w27(self.conv_in, ZQ_PAD), "mid": [res(self.mid.block_1), res(self.mid.block_2)], "up": [],
              "nout": norm(self.norm_out), "out": w27(self.conv_out)}
        for lvl in self.up:
            e = {"res": [res(r) for r in lvl.block]}
            if hasattr(lvl, "upsample"):
                e["up"] = (h(lvl.upsample.conv.weight.detach().permute(0, 2, 3, 1)), h(lvl.upsample.conv.bias),
                           lvl.upsample.compress_time)
            pk["up"].append(e)
        self._packed = pk
        return pk

    # ---- building blocks -----------------------------------------------------------------------------------------------
    def _context(self, key, buf, T, HW, keep):
        """fill the two leading frames of a causal conv's input buffer and (if the next chunk follows) keep its last two"""
        prev = self._cache.pop(key, None)
        if prev is not None:
            buf[:2 * HW].copy_(prev)
        else:
            buf[:HW].copy_(buf[2 * HW:3 * HW])
            buf[HW:2 * HW].copy_(buf[2 * HW:3 * HW])
        if keep:
            self._cache[key] = buf[T * HW:(T + 2) * HW].clone()

    def _norm_into(self, n, x, zq, geom, out):
        T, H, W, Tl, Hl, Wl = geom
        mod = ops.linear(zq, n["w"], n["wb"])                                     # [Tl*Hl*Wl, 2C]: conv_y(zq) | conv_b(zq)
        C = n["C"]
        return ops.groupnorm_mod(x, n["g"], n["b"], mod[:, :C], mod[:, C:], T, H, W, Tl, Hl, Wl, 1e-6, True, out=out)

    def _causal(self, key, x_norm_buf, wb, T, H, W, keep, residual=None, out=None):
        self._context(key, x_norm_buf, T, H * W, keep)
        return ops.conv3d_causal(x_norm_buf, wb[0], T, H, W, wb[1], residual=residual, out=out)

    def _res(self, key, p, x, zq, geom, keep):
        T, H, W = geom[:3]
        HW, Cin, Cout = H * W, x.shape[1], p["c1"][0].shape[0]
        buf = torch.empty(((T + 2) * HW, Cin), dtype=x.dtype, device=x.device)
        self._norm_into(p["n1"], x, zq, geom, buf[2 * HW:])
        h = self._causal(key + ".conv1", buf, p["c1"], T, H, W, keep)
        buf = buf if Cin == Cout else torch.empty(((T + 2) * HW, Cout), dtype=x.dtype, device=x.device)
        self._norm_into(p["n2"], h, zq, geom, buf[2 * HW:])
        skip = ops.linear(x, p["nin"][0], p["nin"][1]) if "nin" in p else x
        return self._causal(key + ".conv2", buf, p["c2"], T, H, W, keep, residual=skip, out=h)

    @staticmethod
    def _upsample(x, T, H, W, up):
        w9, bias, compress_time = up
        C, HW = x.shape[1], H * W
        if compress_time and T > 1:
            src = ([0] + [1 + i // 2 for i in range(2 * (T - 1))]) if T % 2 == 1 else [i // 2 for i in range(2 * T)]
        else:
            src = list(range(T))
        To = len(src)
        big = torch.empty((To * 4 * HW, C), dtype=x.dtype, device=x.device)
        done = {}
        for t, s in enumerate(src):
            dst = big[t * 4 * HW:(t + 1) * 4 * HW]
            if s in done:
                dst.copy_(done[s])
            else:
                done[s] = ops.upsample2x(x[s * HW:(s + 1) * HW], 1, H, W, out=dst)
        return ops.conv2d_3x3(big.view(To, 2 * H, 2 * W, C), w9, bias), To, 2 * H, 2 * W

    # ---- public surface ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, z, clear_fake_cp_cache=True):
        """z (1, z_channels, Tl, Hl, Wl) -> (1, out_ch, T, 8 Hl, 8 Wl) in the model's dtype (cp_enc_dec.py:949-980)."""
        pk = self._packed or self._pack()
        dt = self._dtype()
        B, Cz, Tl, Hl, Wl = z.shape
        assert B == 1 and Cz == self.z_channels
        keep = not clear_fake_cp_cache
        zq = torch.zeros((Tl * Hl * Wl, ZQ_PAD), dtype=dt, device=z.device)
        zq[:, :Cz] = z[0].permute(1, 2, 3, 0).reshape(-1, Cz).to(dt)
        T, H, W = Tl, Hl, Wl
        buf = torch.zeros(((T + 2) * H * W, ZQ_PAD), dtype=dt, device=z.device)
        buf[2 * H * W:] = zq
        h = self._causal("conv_in", buf, pk["in"], T, H, W, keep)
        geom = (T, H, W, Tl, Hl, Wl)
        for i, p in enumerate(pk["mid"]):
            h = self._res(f"mid.{i}", p, h, zq, geom, keep)
        for lvl in reversed(range(self.num_resolutions)):
            e = pk["up"][lvl]
            for j, p in enumerate(e["res"]):
                h = self._res(f"up.{lvl}.{j}", p, h, zq, geom, keep)
            if "up" in e:
                h, T, H, W = self._upsample(h, T, H, W, e["up"])
                geom = (T, H, W, Tl, Hl, Wl)
        C = h.shape[1]
        buf = torch.empty(((T + 2) * H * W, C), dtype=dt, device=z.device)
        self._norm_into(pk["nout"], h, zq, geom, buf[2 * H * W:])
        rgb = torch.empty((T * H * W, 8), dtype=dt, device=z.device)
        self._causal("conv_out", buf, pk["out"], T, H, W, keep, out=rgb[:, :self.out_ch])
        if clear_fake_cp_cache:
            self._cache.clear()
        return ops.tokens_to_nchw5(rgb[:, :self.out_ch], 1, self.out_ch, T, H, W)

    @torch.no_grad()
    def decode_latent(self, latent):
        """the serial chunk loop of sample_sr.py:212-227: latent (1, C, Tl, h, w) -> (1, 3, 4 (Tl - 1) + 1, 8h, 8w);
        the first call takes 3 latent frames, every later one 2, the causal-conv context is carried on the GPU"""
        Tl = latent.shape[2]
        loops = (Tl - 1) // 2
        out = []
        for i in range(loops):
            a, b = (0, 3) if i == 0 else (2 * i + 1, 2 * i + 3)
            out.append(self.forward(latent[:, :, a:b].contiguous(), clear_fake_cp_cache=(i == loops - 1)))
        return torch.cat(out, dim=2)


class ContextParallelEncoder3D(nn.Module):
    """``ContextParallelEncoder3D`` (cp_enc_dec.py:716-836) with the reference's constructor keywords and state-dict keys: the
    encoder that turns the (bicubically pre-upsampled) LQ clip into the latent the DiT is conditioned on
    (diffusion_video.py:279-283).  The whole clip is one call (no chunk protocol on this side): causal 27-tap convs with the first
    frame replicated in front, clip-wide GroupNorm32 + SiLU written straight into the conv's input buffer, DownSample3D =
    frame-pair average (first frame kept for odd T) + stride-2 conv with (0,1,0,1) padding.  ``forward`` returns the moments
    (mean | logvar, 2 z_channels); ``encode`` applies DiagonalGaussianRegularizer (regularizers.py:84-104)."""

    def __init__(self, *, ch=128, out_ch=3, ch_mult=(1, 2, 2, 4), num_res_blocks=3, attn_resolutions=(), dropout=0.0,
                 resamp_with_conv=True, in_channels=3, resolution=256, z_channels=16, double_z=True, pad_mode="first",
                 temporal_compress_times=4, gather_norm=False, **ignore_kwargs):
        super().__init__()
        if len(attn_resolutions) or not resamp_with_conv or not double_z:
            raise NotImplementedError("only the shipped CogVideoX encoder configuration is built")
        assert in_channels <= ZQ_PAD
        self.in_channels, self.z_channels = in_channels, z_channels
        self.num_resolutions = len(ch_mult)
        tlevel = {1: 0, 2: 1, 4: 2, 8: 3}[temporal_compress_times]
        self.conv_in = _CausalConv3d(in_channels, ch, 3)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        for i_level in range(self.num_resolutions):
            down = nn.Module()
            down.block, down.attn = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                down.block.append(_ResnetBlock3DPlain(block_in, block_out))
                block_in = block_out
            if i_level != self.num_resolutions - 1:
                down.downsample = _DownSample3D(block_in, compress_time=i_level < tlevel)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = _ResnetBlock3DPlain(block_in, block_in)
        self.mid.block_2 = _ResnetBlock3DPlain(block_in, block_in)
        self.norm_out = nn.GroupNorm(32, block_in, eps=1e-6, affine=True)
        self.conv_out = _CausalConv3d(block_in, 2 * z_channels, 3)
        self._packed = None

    _dtype = ContextParallelDecoder3D._dtype

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _pack(self):
        dt, dev = self._dtype(), self.conv_in.conv.weight.device

        def h(t):
            return t.detach().to(device=dev, dtype=dt).contiguous()

        def w27(c, pad_in=None):
            w = c.conv.weight.detach().float().permute(0, 2, 3, 4, 1)
            if pad_in:
                w = torch.nn.functional.pad(w, (0, pad_in - w.shape[-1]))
            return h(w), h(c.conv.bias)

        def res(r):
            p = {"n1": (h(r.norm1.weight), h(r.norm1.bias)), "c1": w27(r.conv1), "n2": (h(r.norm2.weight), h(r.norm2.bias)),
                 "c2": w27(r.conv2)}
            if hasattr(r, "nin_shortcut"):
                p["nin"] = (h(r.nin_shortcut.weight.detach()[:, :, 0, 0, 0]), h(r.nin_shortcut.bias))
            return p

        pk = {"in": w27(self.conv_in, ZQ_PAD), "down": [], "mid": [res(self.mid.block_1), res(self.mid.block_2)],
              "nout": (h(self.norm_out.weight), h(self.norm_out.bias)), "out": w27(self.conv_out)}
        for lvl in self.down:
            e = {"res": [res(r) for r in lvl.block]}
            if hasattr(lvl, "downsample"):
                e["down"] = (h(lvl.downsample.conv.weight.detach().permute(0, 2, 3, 1)), h(lvl.downsample.conv.bias),
                             lvl.downsample.compress_time)
            pk["down"].append(e)
        self._packed = pk
        return pk

    @staticmethod
    def _causal(buf, wb, T, H, W, residual=None, out=None):
        HW = H * W
        buf[:HW].copy_(buf[2 * HW:3 * HW])                           # the first frame twice in front (cp_enc_dec.py:268)
        buf[HW:2 * HW].copy_(buf[2 * HW:3 * HW])
        return ops.conv3d_causal(buf, wb[0], T, H, W, wb[1], residual=residual, out=out)

    def _res(self, p, x, T, H, W):
        HW, Cin, Cout = H * W, x.shape[1], p["c1"][0].shape[0]
        buf = torch.empty(((T + 2) * HW, Cin), dtype=x.dtype, device=x.device)
        ops.groupnorm(x, p["n1"][0], p["n1"][1], 1, 1e-6, True, out=buf[2 * HW:])
        h = self._causal(buf, p["c1"], T, H, W)
        buf = buf if Cin == Cout else torch.empty(((T + 2) * HW, Cout), dtype=x.dtype, device=x.device)
        ops.groupnorm(h, p["n2"][0], p["n2"][1], 1, 1e-6, True, out=buf[2 * HW:])
        skip = ops.linear(x, p["nin"][0], p["nin"][1]) if "nin" in p else x
        return self._causal(buf, p["c2"], T, H, W, residual=skip, out=h)

    @torch.no_grad()
    def forward(self, x):
        """x (1, 3, T, H, W) in [-1, 1] -> moments (1, 2 z_channels, (T - 1) / 4 + 1, H / 8, W / 8) in the model's dtype"""
        pk = self._packed or self._pack()
        dt = self._dtype()
        B, Cx, T, H, W = x.shape
        assert B == 1 and Cx == self.in_channels and H % 8 == 0 and W % 8 == 0
        buf = torch.zeros(((T + 2) * H * W, ZQ_PAD), dtype=dt, device=x.device)
        buf[2 * H * W:, :Cx] = x[0].permute(1, 2, 3, 0).reshape(-1, Cx).to(dt)
        h = self._causal(buf, pk["in"], T, H, W)
        del buf
        for e in pk["down"]:
            for p in e["res"]:
                h = self._res(p, h, T, H, W)
            if "down" in e:
                w9, bias, compress_time = e["down"]
                if compress_time and T > 1:
                    h = ops.time_avgpool2(h, T, H * W)
                    T = (T + 1) // 2 if T % 2 else T // 2
                h, H, W = ops.conv2d_3x3_s2p(h.view(T, H, W, h.shape[1]), w9, bias, pad=(0, 1, 0, 1))
        for p in pk["mid"]:
            h = self._res(p, h, T, H, W)
        C = h.shape[1]
        buf = torch.empty(((T + 2) * H * W, C), dtype=dt, device=x.device)
        ops.groupnorm(h, pk["nout"][0], pk["nout"][1], 1, 1e-6, True, out=buf[2 * H * W:])
        m = self._causal(buf, pk["out"], T, H, W)
        return ops.tokens_to_nchw5(m, 1, 2 * self.z_channels, T, H, W)

    @torch.no_grad()
    def encode(self, x, sample=True):
        """``VideoAutoencodingEngine.encode`` (autoencoder.py:218-230): encoder + DiagonalGaussianRegularizer -> z (1, 16, Tl, h, w)"""
        mean, logvar = torch.chunk(self.forward(x), 2, dim=1)
        if not sample:
            return mean
        std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))
        return (mean.float() + std.float() * torch.randn_like(mean, dtype=torch.float32)).to(mean.dtype)      # noise drawn in fp32
