"""CUDA-graphed denoiser call (SURVEY K13 / VERDICT r1 item 9).

One solver step of the chunked path launches ~3 300 kernels through the C ABI.  At BASELINE config 2 they average 400 us, the
GPU queue never drains and the host is invisible; at small latents (BASELINE config 1, previews, the 'fast' schedule on short
clips) the same 3 300 launches take longer to ISSUE (~10 us each: ctypes call, tensor-map encode, output allocation) than to run.
``GraphedCFGPair`` captures ``ControlledV2VUNet.forward_cfg_pair`` -- both CFG branches of a step, the whole UNet + ControlNet --
into one CUDA graph per input shape and replays it with the step's (x, t, hint, y) copied into static buffers.  Everything the
kernels need is baked at capture: device pointers and TMA tensor maps are kernel arguments, outputs live in the graph's memory
pool.  The result is bit-identical to the eager call (same kernels, same order).

    model = GraphedCFGPair(generator)            # drop-in for `model=` of GaussianDiffusion.sample_sr / denoise
"""
import torch


class _Captured:
    __slots__ = ("graph", "x", "t", "y0", "y1", "hint", "out")


class GraphedCFGPair(torch.nn.Module):
    """wraps a ControlledV2VUNet; ``forward_cfg_pair`` is served from CUDA graphs (one per chunk shape), ``forward`` stays eager"""

    def __init__(self, model, max_graphs=4, warmup=2):
        super().__init__()
        self.model = model
        self._graphs = {}
        self._max, self._warmup = max_graphs, warmup
        self.replays = 0

    def forward(self, *a, **k):
        return self.model(*a, **k)

    def half(self):
        self.model = self.model.half()
        return self

    @torch.no_grad()
    def _capture(self, x, t, y_pair, hint):
        c = _Captured()
        c.x, c.t, c.hint = x.clone(), t.clone(), hint.clone()
        c.y0, c.y1 = y_pair[0].clone(), y_pair[1].clone()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):                                   # warm-up off the default stream: lazy init, weight packing
            for _ in range(self._warmup):
                self.model.forward_cfg_pair(c.x, c.t, (c.y0, c.y1), hint=c.hint)
        torch.cuda.current_stream(x.device).wait_stream(side)
        torch.cuda.synchronize(x.device)
        c.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(c.graph):
            c.out = self.model.forward_cfg_pair(c.x, c.t, (c.y0, c.y1), hint=c.hint)
        return c

    @torch.no_grad()
    def forward_cfg_pair(self, x, t, y_pair, hint=None, hint_chunk=None, variant_info=None):
        if hint_chunk is not None:
            hint = hint_chunk
        key = (tuple(x.shape), x.dtype, tuple(hint.shape), hint.dtype, tuple(y_pair[0].shape), y_pair[0].dtype, t.dtype, x.device.index)
        c = self._graphs.get(key)
        if c is None:
            if len(self._graphs) >= self._max:                          # unusual shape mix: do not hoard graph pools
                return self.model.forward_cfg_pair(x, t, y_pair, hint=hint)
            c = self._graphs[key] = self._capture(x, t, y_pair, hint)
        c.x.copy_(x)
        c.t.copy_(t)
        c.hint.copy_(hint)
        c.y0.copy_(y_pair[0])
        c.y1.copy_(y_pair[1])
        c.graph.replay()
        self.replays += 1
        return tuple(o.clone() for o in c.out)                          # the static outputs are overwritten by the next replay
