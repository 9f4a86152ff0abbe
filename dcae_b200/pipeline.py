"""Host-resident streaming front end of the slice loop: the caller hands pinned HOST tensors
(y, latent_scales, latent_means) and receives HOST results; H2D copies, compute and D2H copies of
consecutive batches overlap on three CUDA streams with double buffering.  This is the call a user of the
reference makes when latents live in host memory (e.g. between the hyper-decoder on another device and the
rANS coder on the CPU, dcae.py:742-756); `bench.py` measures its `e2e` number through it."""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Sequence

import torch

from .entropy_model import EntropySliceLoop

OUT_KEYS = ("y_hat", "means", "scales", "likelihoods")


class HostPipeline:
    """mode "forward": the four fp32 tensors of DCAE.forward (+ int32 symbols / indexes with want_symbols) come back.
    mode "compress": what DCAE.compress hands to the range coder comes back -- int16 symbols, uint8 indexes in coder
    order (dcae.py:742-743; `dcae_pack_symbols`), the overflow count and the bpp numerator; y_hat stays on the device
    for g_s (`dev_out[k]["y_hat"]`): 3 bytes per element over PCIe instead of 16."""

    def __init__(self, engine: EntropySliceLoop, B: int, h: int, w: int, depth: int = 2, want_symbols: bool = False,
                 mode: str = "forward"):
        if mode not in ("forward", "compress"):
            raise ValueError(mode)
        self.mode = mode
        want_symbols = want_symbols or mode == "compress"
        self.eng, self.depth, self.want_symbols = engine, depth, want_symbols
        dev = engine.device
        self.s_h2d, self.s_cmp, self.s_d2h = (torch.cuda.Stream(dev) for _ in range(3))
        shape = (B, 320, h, w)
        self.dev_in = [[torch.zeros(shape, device=dev) for _ in range(3)] for _ in range(depth)]     # zeros: the warm call below runs on them
        self.dev_out: List[Dict[str, torch.Tensor]] = []
        self.host_out: List[Dict[str, torch.Tensor]] = []
        keys = OUT_KEYS + (("symbols", "indexes") if want_symbols else ())
        if mode == "compress":
            keys = ("symbols16", "indexes8", "overflow", "log2_lik_sum")
        with torch.cuda.stream(self.s_cmp):
            for k in range(depth):      # one warm call per slot creates the output buffers (and the plan)
                o = engine.forward(*self.dev_in[k], want_symbols=want_symbols)
                if mode == "compress":
                    n = o["symbols"].numel()
                    o["symbols16"] = torch.empty(n, dtype=torch.int16, device=dev)
                    o["indexes8"] = torch.empty(n, dtype=torch.uint8, device=dev)
                    o["overflow"] = torch.zeros(1, dtype=torch.int64, device=dev)
                self.dev_out.append(o)
                self.host_out.append({key: torch.empty(o[key].shape, dtype=o[key].dtype).pin_memory() for key in keys})
        self.s_cmp.synchronize()
        self.ev_in = [torch.cuda.Event() for _ in range(depth)]       # inputs of slot k landed
        self.ev_cmp = [torch.cuda.Event() for _ in range(depth)]      # compute of slot k finished
        self.ev_out = [torch.cuda.Event() for _ in range(depth)]      # outputs of slot k are on the host
        self.h2d_bytes = 3 * self.dev_in[0][0].numel() * 4
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.host_out[0].values())

    def run(self, batches: Iterable[Sequence[torch.Tensor]]) -> Iterator[Dict[str, torch.Tensor]]:
        """batches: iterable of (y, latent_scales, latent_means) pinned host tensors.  Yields, in order, dicts
        of pinned host tensors; a yielded dict is valid until `depth` more batches have been consumed."""
        pending: List[int] = []
        for i, batch in enumerate(batches):
            k = i % self.depth
            if i >= self.depth:                         # slot k is being reused: hand out its previous result
                self.ev_out[k].synchronize()
                pending.pop(0)
                yield self.host_out[k]
            with torch.cuda.stream(self.s_h2d):
                self.s_h2d.wait_event(self.ev_cmp[k])   # the previous compute on this slot has consumed its inputs
                for dst, src in zip(self.dev_in[k], batch):
                    dst.copy_(src, non_blocking=True)
                self.ev_in[k].record(self.s_h2d)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(self.ev_in[k])
                self.s_cmp.wait_event(self.ev_out[k])   # the previous outputs of this slot have left the device
                o = self.eng.forward(*self.dev_in[k], want_symbols=self.want_symbols, out=self.dev_out[k])
                if self.mode == "compress":
                    from . import _lib
                    with torch.cuda.device(self.eng.device):
                        _lib.check(self.eng.lib.dcae_pack_symbols(o["symbols"].data_ptr(), o["indexes"].data_ptr(), o["symbols"].numel(),
                                                                  o["symbols16"].data_ptr(), o["indexes8"].data_ptr(), o["overflow"].data_ptr(),
                                                                  self.s_cmp.cuda_stream), "dcae_pack_symbols")
                self.ev_cmp[k].record(self.s_cmp)
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(self.ev_cmp[k])
                for key, dst in self.host_out[k].items():
                    dst.copy_(self.dev_out[k][key], non_blocking=True)
                self.ev_out[k].record(self.s_d2h)
            pending.append(k)
        for k in pending:
            self.ev_out[k].synchronize()
            yield self.host_out[k]
