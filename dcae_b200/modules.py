"""Module-level drop-ins for the reference's own loop text.

`DCAE.forward / compress / decompress` (dcae.py:638-670, 713-753, 878-906) call four kinds of sub-modules inside
the slice loop: `self.dt_cross_attention[i](query, dt)`, `self.cc_mean_transforms[i](support)`,
`self.cc_scale_transforms[i](support)`, `self.lrp_transforms[i](lrp_support)` and `self.gaussian_conditional`.
`accelerate(net)` redirects exactly those calls of a reference `DCAE` instance to libdcae_b200.so, so the
reference's own Python loop runs unchanged on the CUDA kernels.  (The whole-loop object, `EntropySliceLoop`, is
faster: it keeps everything token-major between modules and fuses the first conv layer of the three stacks; this
file is for maintainers who want to keep `dcae.py` as it is.)

The reference modules stay where they are: only their `forward` is overridden on the instance.  Their parameters
therefore remain in `net.parameters()` / `net.state_dict()` under the reference's keys, `net.load_state_dict(ckpt)`
keeps working, and the packed device weights are refreshed automatically when any hot-path parameter has changed
(tensor version counters), e.g. after `load_state_dict` or an optimizer step.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _lib
from .entropy_model import EntropySliceLoop
from .gaussian_conditional import GaussianConditional
from .params import DICT_DIM, DICT_NUM, M_LATENT, NUM_SLICES, SLICE_CH

HOT_PREFIXES = ("dt", "dt_cross_attention", "cc_mean_transforms", "cc_scale_transforms", "lrp_transforms")


def _no_grad_path(*tensors) -> None:
    """The stand-alone module classes below are inference kernels: refuse to run silently inside an autograd graph
    (`accelerate(net)` on a model with parameters differentiates through `dcae_b200.training.ModuleFunction`; the
    whole-loop training path is `dcae_b200.training.SliceLoopFunction` / `EntropyModel`)."""
    if torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors):
        raise _lib.DcaeError("this dcae_b200 module is forward-only and an input requires grad: wrap the call in "
                             "torch.no_grad(), or use accelerate(net) / dcae_b200.training for the training step")


class _LoopModule(torch.nn.Module):
    def __init__(self, loop: EntropySliceLoop, i: int):
        super().__init__()
        self._loop, self.i = [loop], i          # in a list: the engine is not a sub-module / parameter owner

    @property
    def loop(self) -> EntropySliceLoop:
        return self._loop[0]


class DictCrossAttention(_LoopModule):
    """Stand-alone `dt_cross_attention[i]`: forward(x [B, 640 + 64 i, h, w], dt) -> [B, 320, h, w]  (dcae.py:479-509).
    The dictionary is the one the engine was built with; a `dt` argument is checked against it."""

    def forward(self, x: torch.Tensor, dt: Optional[torch.Tensor] = None) -> torch.Tensor:
        _no_grad_path(x, dt)
        _check_dt(self.loop, dt)
        return self.loop.module_dca(self.i, x)


class ConvStack(_LoopModule):
    """Stand-alone `cc_mean_transforms[i]` (which=0), `cc_scale_transforms[i]` (1), `lrp_transforms[i]` (2)."""

    def __init__(self, loop: EntropySliceLoop, i: int, which: int):
        super().__init__(loop, i)
        self.which = which

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        _no_grad_path(x)
        return self.loop.module_conv(self.i, self.which, x)


def _check_dt(loop: EntropySliceLoop, dt: Optional[torch.Tensor], strict: bool = False) -> None:
    """`dt` is `self.dt.repeat([b, 1, 1])` in the reference (dcae.py:625): K and V come from the engine's packed copy of
    the dictionary, so a different `dt` must not be ignored silently.  Shape is checked on every call, values on the
    first call (and on every call with strict=True; the comparison synchronises the stream)."""
    if dt is None:
        return
    if dt.shape[-2:] != (DICT_NUM, DICT_DIM):
        raise ValueError(f"dt: expected [..., {DICT_NUM}, {DICT_DIM}], got {tuple(dt.shape)}")
    if strict or not getattr(loop, "_dt_checked", False):
        ref = loop.dictionary
        d0 = dt.reshape(-1, DICT_NUM, DICT_DIM)[0].to(ref.device, torch.float32)
        if not torch.equal(d0, ref):
            raise ValueError("dt differs from the dictionary packed into this engine (K = k(LN(dt)), V = LN(dt) are "
                             "computed once per weight load): call engine.refresh(net.state_dict()) / accelerate(net) again")
        loop._dt_checked = True


class Accelerated:
    """Handle returned by `accelerate`: the engine plus the bookkeeping that keeps it in step with `net`."""

    def __init__(self, net: torch.nn.Module, loop: EntropySliceLoop, tracked: Dict[str, torch.Tensor], strict_dt: bool):
        self.net, self.loop, self._tracked, self.strict_dt = [net], loop, tracked, strict_dt
        self._sig = self._signature()
        self._orig_forward = {}

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in self._tracked.values())

    def sync(self) -> None:
        """Repack the device weights if a hot-path parameter of `net` changed since they were packed."""
        if self._tracked and self._signature() != self._sig:
            self.loop.refresh({k: v.detach() for k, v in self._tracked.items()})
            self._sig = self._signature()

    def restore(self) -> None:
        """Undo `accelerate`: the reference modules run their own `forward` again."""
        for mod in self._orig_forward:
            mod.__dict__.pop("forward", None)
        self._orig_forward.clear()

    def __getattr__(self, name):          # the handle doubles as the whole-loop object
        return getattr(self.loop, name)


def accelerate(net: torch.nn.Module, device="cuda:0", math: str = "f16x3", state_dict=None, strict_dt: bool = False) -> Accelerated:
    """Redirect the hot-path sub-modules of a reference `DCAE` instance to the CUDA library, in place.
    `state_dict`: take the weights from here instead of `net`'s own parameters (same keys; no change tracking then).
    Returns a handle (`.loop` is the `EntropySliceLoop`; `.restore()` undoes the redirection)."""
    if state_dict is None:
        tracked = {k: v for k, v in list(net.named_parameters()) + list(net.named_buffers()) if k.split(".")[0] in HOT_PREFIXES}
        hot = {k: v.detach() for k, v in tracked.items()}
    else:
        tracked = {}
        hot = {k: v.detach() for k, v in state_dict.items() if k.split(".")[0] in HOT_PREFIXES}
    table = getattr(net.gaussian_conditional, "scale_table", None)
    loop = EntropySliceLoop(hot, device=device, math=math, lanes=1,
                            scale_table=table if table is not None and table.numel() else None)
    handle = Accelerated(net, loop, tracked, strict_dt)

    def patch(mod, fn):
        handle._orig_forward[mod] = True
        mod.forward = fn                      # instance attribute: nn.Module.__call__ picks it up before the class method

    def needs_grad(prefix, *ins):
        """Under autograd with something to differentiate: the call becomes a `training.ModuleFunction` node."""
        if not torch.is_grad_enabled():
            return None
        sub = {k[len(prefix):]: v for k, v in tracked.items() if k.startswith(prefix)}
        if any(isinstance(t, torch.Tensor) and t.requires_grad for t in ins) or any(v.requires_grad for v in sub.values()):
            if not sub:
                _no_grad_path(*ins)          # built from a detached state dict: nothing to differentiate into
            return sub
        return None

    from . import torch_graph

    for i in range(NUM_SLICES):
        def dca(x, dt=None, i=i):
            handle.sync()
            _check_dt(loop, dt, strict_dt)
            sub = needs_grad(f"dt_cross_attention.{i}.", x, dt)
            if sub is None:
                return loop.module_dca(i, x)
            from .training import ModuleFunction
            if dt is None:
                dt = tracked["dt"].unsqueeze(0)
            # dt arrives as self.dt.repeat([b, 1, 1]) (dcae.py:625): K / V are batch invariant, image 0's copy carries the gradient
            graph = lambda ins, P: torch_graph.dictionary_cross_attention(ins[0], ins[1].reshape(-1, DICT_NUM, DICT_DIM)[0], P)   # noqa: E731
            return ModuleFunction.apply(lambda xx, dd: loop.module_dca(i, xx), graph, list(sub), 2, x, dt, *sub.values())

        def conv(x, i=i, which=0):
            handle.sync()
            name = ("cc_mean_transforms", "cc_scale_transforms", "lrp_transforms")[which]
            sub = needs_grad(f"{name}.{i}.", x)
            if sub is None:
                return loop.module_conv(i, which, x)
            from .training import ModuleFunction
            graph = lambda ins, P: torch_graph.conv_stack(ins[0], P)      # noqa: E731
            return ModuleFunction.apply(lambda xx: loop.module_conv(i, which, xx), graph, list(sub), 1, x, *sub.values())

        patch(net.dt_cross_attention[i], dca)
        for which, name in enumerate(("cc_mean_transforms", "cc_scale_transforms", "lrp_transforms")):
            patch(getattr(net, name)[i], lambda x, i=i, which=which: conv(x, i, which))
    old = net.gaussian_conditional
    fast = GaussianConditional(None).to(device)
    fast.load_state_dict(old.state_dict(), strict=False)     # the four table buffers are resized to the incoming shapes
    if fast.scale_table.numel() == 0:
        fast.update_scale_table(loop.scale_table)
    fast.train(old.training)
    net.gaussian_conditional = fast
    return handle


__all__ = ["DictCrossAttention", "ConvStack", "Accelerated", "accelerate", "HOT_PREFIXES", "M_LATENT", "SLICE_CH"]
