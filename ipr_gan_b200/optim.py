"""``FlatAdam``: ``torch.optim.Adam`` whose ``step()`` is one sm_100a launch over the network's flat arenas
(csrc/optim.cu).  It IS a ``torch.optim.Adam`` (same ``param_groups`` / ``state_dict`` format, so checkpoints
written by the reference load and vice versa), only the arithmetic moved.

With ``torch.distributed`` initialised (one process per GPU) ``step()`` first all-reduces (sum) the gradient arena over
NCCL -- one call per network per step, carrying the step's loss slots along -- and the Adam launch applies the
1/world factor: the reference's ``nn.DataParallel`` gradient reduction (experiments/base.py:36-39) without the
per-forward parameter broadcast.
"""
import ctypes

import torch

from . import dist, flat
from ._lib import check, lib


class FlatAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, **other):
        if amsgrad:
            raise NotImplementedError("FlatAdam: amsgrad is not on the IPR-GAN path")
        # options that do not change the arithmetic are accepted, anything else is refused (never silently dropped)
        harmless = {"capturable": None, "foreach": None, "fused": None, "differentiable": False, "maximize": False}
        for k, v in other.items():
            if k not in harmless or (harmless[k] is not None and v != harmless[k]):
                raise NotImplementedError("FlatAdam: option %s=%r is not on the IPR-GAN path" % (k, v))
        params = list(params)
        super().__init__(params, lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        if len(self.param_groups) != 1:
            raise NotImplementedError("FlatAdam: a single parameter group is expected")
        self.arena = flat.arena_for(params)
        dev = self.arena.param.device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam needs CUDA parameters (no CPU fallback)")
        self._m = torch.zeros_like(self.arena.param)
        self._v = torch.zeros_like(self.arena.param)
        self._step = torch.zeros((), device=dev, dtype=torch.float32)
        self._ticket = torch.zeros(1, device=dev, dtype=torch.int32)
        self._bind_state()

    def _bind_state(self):
        for p, o in zip(self.arena.params, self.arena.offsets):
            n = p.numel()
            self.state[p] = {"step": self._step, "exp_avg": self._m[o:o + n].view(p.shape),
                             "exp_avg_sq": self._v[o:o + n].view(p.shape)}

    def zero_grad(self, set_to_none=True):
        """Zeroes the gradient arena (the ``.grad`` views stay bound).  Unlike ``torch.optim.Adam`` a parameter that
        receives no gradient in a step is therefore updated with a zero gradient (its moments decay) instead of being
        skipped; every parameter of G and D receives a gradient in every IPR-GAN step, so the two agree on the path."""
        self.arena.zero_grad()

    def state_dict(self):
        """``torch.optim.Adam`` format with one independent CPU ``step`` tensor per parameter: the single device-side
        counter all parameters share here must not leave as one aliased tensor (a stock Adam resuming from it would
        advance it once per parameter)."""
        sd = super().state_dict()
        step = self._step.detach().to("cpu", copy=True)
        # the packed per-parameter dicts are the live ``self.state`` entries: re-wrap, never mutate them
        sd["state"] = {k: dict(st, step=step.clone()) for k, st in sd["state"].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        with torch.no_grad():
            step = None
            for p, o in zip(self.arena.params, self.arena.offsets):
                st = self.state.get(p)
                if st:
                    n = p.numel()
                    self._m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                    self._v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                    step = st["step"]
            if step is not None:
                self._step.fill_(float(step))
        self._bind_state()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        arena = self.arena
        if not arena.valid():
            raise RuntimeError("FlatAdam: the parameters were moved out of their arena (module.to() / a new p.data "
                               "after the optimizer was built); rebuild the optimizer")
        arena.bind_grads()
        # ONE summing all-reduce per network per step over [gradients | metric slots] (no-op for a single process);
        # the mean's 1/world is folded into the Adam launch, which also clears the arena as it consumes it
        dist.allreduce_sum_(arena.reduce_view)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().ipr_adam_flat_f32(ctypes.c_void_p(arena.param.data_ptr()), ctypes.c_void_p(arena.grad.data_ptr()),
                                      ctypes.c_void_p(self._m.data_ptr()), ctypes.c_void_p(self._v.data_ptr()),
                                      arena.numel, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]),
                                      float(g["eps"]), float(g["weight_decay"]), 1.0 / dist.world(), 1,
                                      ctypes.c_void_p(self._step.data_ptr()), ctypes.c_void_p(self._ticket.data_ptr()),
                                      st), "ipr_adam_flat_f32")
        arena.version += 1
        arena.clean = True
        return loss
