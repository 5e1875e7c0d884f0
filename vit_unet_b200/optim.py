"""Fused AdamW on the CUDA path (the reference trains with
``torch.optim.AdamW(model.parameters(), lr)`` -- run_denoising.py:81).

``FusedAdamW`` is a drop-in ``torch.optim.Optimizer`` with torch.optim.AdamW's update rule, executed by the
``vu_adamw`` kernel.  Two paths, both on the GPU:

* per-tensor: one launch per parameter (works for any parameter list);
* flat: after ``FusedAdamW.flatten(model)`` the parameters of a vit_unet_b200 model are views of ONE flat buffer laid
  out exactly like the flat gradient buffer backward writes (forward-execution order), so the whole step is a
  single launch over ~37 M elements when the gradients still alias that buffer.
"""
from __future__ import annotations

import torch

from . import ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._flat = None          # (model, flat_param, flat_m, flat_v, step)
        self._warned = False

    # ------------------------------------------------------------------------------------------ flat path
    def flatten(self, model) -> "FusedAdamW":
        """Re-home the model's parameters into one flat fp32 buffer (same offsets as the flat gradient buffer)."""
        pd = dict(model.named_parameters())
        dev = next(iter(pd.values())).device
        flat = torch.zeros(model._flat_numel, dtype=torch.float32, device=dev)
        for name, off in zip(model._param_names, model._flat_offsets):
            p = pd[name]
            view = flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        step = torch.zeros((), dtype=torch.int64)          # ONE shared step counter (host tensor, updated in place)
        for name, off in zip(model._param_names, model._flat_offsets):
            p = pd[name]
            old = self.state.get(p, {})
            st = dict(step=step, m=m[off:off + p.numel()].view(p.shape), v=v[off:off + p.numel()].view(p.shape))
            if old:                                         # flatten() after some per-tensor steps: keep the moments
                st["m"].copy_(old["m"]); st["v"].copy_(old["v"]); step.fill_(int(old["step"]))
            self.state[p] = st
        # self.state[p]['m'/'v'] are VIEWS of the flat moment buffers and every parameter shares one step tensor, so
        # state_dict()/load_state_dict() round-trip the flat path and the per-tensor fall-back continues from the
        # same moments instead of restarting at zero.
        self._flat = dict(model=model, p=flat, m=m, v=v, step=step)
        self._warned = False
        return self

    def load_state_dict(self, state_dict):
        """torch's loader replaces the per-parameter state tensors; copy them back INTO the flat buffers."""
        if self._flat is None:
            return super().load_state_dict(state_dict)
        views = {p: dict(st) for p, st in self.state.items()}
        super().load_state_dict(state_dict)
        step = self._flat["step"]
        for p, old in views.items():
            new = self.state.get(p)
            if new:
                old["m"].copy_(new["m"]); old["v"].copy_(new["v"]); step.fill_(int(new["step"]))
            self.state[p] = old

    def _flat_grads_alias(self) -> bool:
        f = self._flat
        g = f["model"].flat_grad()
        if g is None or g.numel() != f["p"].numel():
            return False
        pd = dict(f["model"].named_parameters())
        base = g.data_ptr()
        for name, off in zip(f["model"]._param_names, f["model"]._flat_offsets):
            p = pd[name]
            if p.grad is None or p.grad.data_ptr() != base + 4 * off or p.data.data_ptr() != f["p"].data_ptr() + 4 * off:
                return False
        return True

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self._flat is not None and len(self.param_groups) == 1 and self._flat_grads_alias():
            f, grp = self._flat, self.param_groups[0]
            f["step"] += 1
            ops.adamw(f["p"], f["model"].flat_grad(), f["m"], f["v"], grp["lr"], grp["betas"][0], grp["betas"][1],
                      grp["eps"], grp["weight_decay"], int(f["step"]))
            return loss
        if self._flat is not None and not self._warned:
            import warnings
            warnings.warn("FusedAdamW: gradients no longer alias the model's flat gradient buffer (zero_grad("
                          "set_to_none=False), gradient accumulation or model.to() after flatten()); using one launch "
                          "per parameter on the same moment buffers", stacklevel=2)
            self._warned = True
        shared_step_done = False
        for grp in self.param_groups:
            for p in grp["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("FusedAdamW runs on CUDA tensors only (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.int64)
                    st["m"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["v"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if self._flat is not None and st["step"] is self._flat["step"]:
                    if not shared_step_done:                 # one shared counter: advance it once per step()
                        st["step"] += 1
                        shared_step_done = True
                else:
                    st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                ops.adamw(p.data, g, st["m"], st["v"], grp["lr"], grp["betas"][0], grp["betas"][1], grp["eps"],
                          grp["weight_decay"], int(st["step"]))
        return loss
