"""FusedAdam: drop-in for the ``torch.optim.Adam(optimizer_config)`` the host code builds
(src/Mapper.py:111-139,358-364; src/Tracker.py:324-329): same constructor shape (list of param groups with
``lr`` / ``betas`` / ``eps``), same update rule, ONE kernel launch per step over every parameter
(``usl_adam_step``), optionally clearing the gradients in the same pass.

Like torch, the step count lives in ``state[p]['step']`` PER PARAMETER: a parameter whose ``.grad`` is None during early
steps (or that joins later through add_param_group) gets its own bias correction.  ``state_dict`` / ``load_state_dict``
use torch.optim's layout ({'state': {index: {...}}, 'param_groups': [...]}), so checkpoints interchange with
torch.optim.Adam.  With zero_grad_in_step=True the gradients are cleared inside the update kernel -- do not combine it with a
driver that zero-fills the same buffer itself (MappingStep clears its flat gradient buffer at the start of run())."""
from ctypes import byref

import torch

from . import _lib as L


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, zero_grad_in_step=False):
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps)
        self.param_groups = []
        self.state = {}
        self.zero_grad_in_step = zero_grad_in_step
        self._step_dev = None
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        for g in groups:
            self.add_param_group(g)

    def add_param_group(self, g):
        ps = g["params"]
        ps = [ps] if isinstance(ps, torch.Tensor) else list(ps)
        self.param_groups.append({"params": ps, "lr": g.get("lr", self.defaults["lr"]), "betas": tuple(g.get("betas", self.defaults["betas"])),
                                  "eps": g.get("eps", self.defaults["eps"])})
        n = sum(len(gr["params"]) for gr in self.param_groups)
        if n > L.ADAM_MAX_GROUPS:
            raise ValueError(f"FusedAdam: {n} parameter tensors, one launch covers at most {L.ADAM_MAX_GROUPS}")

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
                    st = self.state[p] = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
                out.append((p, g, st))
        return out

    def enable_graph_step_counter(self, device):
        """Keep ONE step count on the device so ``step()`` can be captured in / replayed from a CUDA graph (every parameter
        must then take part in every step, as in the tracker's and the mapper's loops)."""
        cur = max([st["step"] for st in self.state.values()], default=0)
        self._step_dev = torch.full((1,), cur, device=device, dtype=torch.int64)

    @torch.no_grad()
    def step(self):
        items = self._flat()
        if not items:
            return
        if self._step_dev is not None:
            self._step_dev += 1
        arr = (L.AdamGroup * len(items))()
        for a, (p, g, st) in zip(arr, items):
            st["step"] += 1
            a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            a.n, a.lr, a.beta1, a.beta2, a.eps = p.numel(), g["lr"], g["betas"][0], g["betas"][1], g["eps"]
            a.step = 0 if self._step_dev is not None else st["step"]
        L.call("usl_adam_step", arr, len(items), 1, L.ptr(self._step_dev), int(self.zero_grad_in_step), L.stream())

    @torch.no_grad()
    def reset_state(self):
        """Back to a freshly constructed optimiser (zero moments, step 0) without reallocating: what the host code gets by
        building a new torch.optim.Adam per tracked / mapped frame (Tracker.py:324-329, Mapper.py:364), but with stable
        state pointers, so a CUDA graph that captured step() stays valid across frames."""
        for st in self.state.values():
            st["step"] = 0
            st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
        if self._step_dev is not None:
            self._step_dev.zero_()

    def zero_grad(self, set_to_none=False):
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()

    # ---- torch.optim-compatible checkpointing -----------------------------------------------------------------------
    def state_dict(self):
        index, groups, k = {}, [], 0
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                index[p] = k; ids.append(k); k += 1
            groups.append({"lr": g["lr"], "betas": g["betas"], "eps": g["eps"], "params": ids})
        state = {index[p]: {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"].clone(), "exp_avg_sq": st["exp_avg_sq"].clone()}
                 for p, st in self.state.items() if p in index}
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        flat = [p for g in self.param_groups for p in g["params"]]
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            g["lr"], g["betas"], g["eps"] = sg["lr"], tuple(sg["betas"]), sg["eps"]
        self.state = {}
        for k, st in sd["state"].items():
            p = flat[int(k)]
            self.state[p] = {"step": int(st["step"]), "exp_avg": st["exp_avg"].to(p.device, torch.float32).clone(),
                             "exp_avg_sq": st["exp_avg_sq"].to(p.device, torch.float32).clone()}
        if self._step_dev is not None:
            self._step_dev.fill_(max([s["step"] for s in self.state.values()], default=0))
