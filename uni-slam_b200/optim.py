"""FusedAdam: drop-in for the ``torch.optim.Adam(optimizer_config)`` the host code builds
(src/Mapper.py:111-139,358-364; src/Tracker.py:324-329): same constructor shape (list of param groups with
``lr`` / ``betas`` / ``eps``), same update rule, ONE kernel launch per step over every parameter
(``usl_adam_step``), optionally clearing the gradients in the same pass."""
from ctypes import byref

import torch

from . import _lib as L


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, zero_grad_in_step=False):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups = []
        for g in groups:
            ps = g["params"]
            ps = [ps] if isinstance(ps, torch.Tensor) else list(ps)
            self.param_groups.append({"params": ps, "lr": g.get("lr", lr), "betas": tuple(g.get("betas", betas)), "eps": g.get("eps", eps)})
        self.state = {}
        self.zero_grad_in_step = zero_grad_in_step
        self._step = 0
        self._step_dev = None

    def _flat(self):
        out = []
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("FusedAdam: parameters and gradients must be contiguous fp32 CUDA tensors")
                st = self.state.get(p)
                if st is None:
                    st = self.state[p] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
                out.append((p, g, st))
        return out

    def enable_graph_step_counter(self, device):
        """Keep the step count on the device so ``step()`` can be captured in / replayed from a CUDA graph."""
        self._step_dev = torch.full((1,), self._step, device=device, dtype=torch.int64)

    @torch.no_grad()
    def step(self):
        items = self._flat()
        if not items:
            return
        if len(items) > L.ADAM_MAX_GROUPS:
            raise RuntimeError(f"FusedAdam: at most {L.ADAM_MAX_GROUPS} tensors per step")
        self._step += 1
        if self._step_dev is not None:
            self._step_dev += 1
        arr = (L.AdamGroup * len(items))()
        for a, (p, g, st) in zip(arr, items):
            a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            a.n, a.lr, a.beta1, a.beta2, a.eps = p.numel(), g["lr"], g["betas"][0], g["betas"][1], g["eps"]
        L.call("usl_adam_step", arr, len(items), self._step, L.ptr(self._step_dev), int(self.zero_grad_in_step), L.stream())

    def zero_grad(self, set_to_none=False):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()
