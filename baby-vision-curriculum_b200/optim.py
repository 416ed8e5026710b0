"""FusedSGD -- torch.optim.SGD's interface (the optimizer of pretrain_videomae.py:187-189: SGD, nesterov, momentum
0.9, lr 0.1, wd 0) on one libbvc.so multi-tensor kernel (csrc/optim.cu, bvc_sgd_step).

Drop-in for the reference's optimizer line; state-dict layout is torch.optim.SGD's (`momentum_buffer` per parameter).
Works under torch.amp.GradScaler through the `_step_supports_amp_scaling` protocol: GradScaler hands over `grad_scale`
/ `found_inf` device tensors, the kernel unscales (writing the unscaled gradient back, which is what the reference's
grad logger reads after scaler.step -- loggingtools.py:107-118) and skips the update on overflow, all on the device.

`shadow_from=model` (a bvc_b200.VideoMAEForPreTraining, possibly DDP-wrapped) additionally refreshes the model's bf16
operand copies of the weights inside the same pass, so the next forward does not re-cast them.
No CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


class FusedSGD(torch.optim.Optimizer):
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, *,
                 maximize=False, shadow_from=None):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if momentum < 0.0:
            raise ValueError(f"Invalid momentum value: {momentum}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")  # torch/optim/sgd.py
        if maximize:
            raise NotImplementedError("maximize=True is not on the reference's path")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay,
                                      nesterov=nesterov))
        self._shadow_from = shadow_from
        self._tables = {}        # group index -> (key, device table, n_entries, total elements, bytes per element)
        self.table_builds = 0    # how often a pointer table had to be rebuilt (stable pointers -> stays small)

    # ------------------------------------------------------------------------------------------------------------
    def _shadow_model(self):
        m = self._shadow_from
        if m is None:
            return None
        return getattr(m, "module", m)

    def _table(self, gi, group, plist):
        model = self._shadow_model()
        shadows = model.weight_shadows() if model is not None else {}  # {} unless the copies match the weights now
        if shadows:
            self._shadows_used = True
        mom = group["momentum"] != 0
        rows, key = [], []
        for p in plist:
            g = p.grad
            st = self.state[p]
            uninit = 0
            if mom and "momentum_buffer" not in st:
                st["momentum_buffer"] = torch.empty_like(p, memory_format=torch.preserve_format)
                st["_bvc_uninit"] = True
            if mom and st.get("_bvc_uninit", False):
                uninit = 1
            sh = shadows.get(id(p), (0, 0))
            rows.append((p.data_ptr(), g.data_ptr(), st["momentum_buffer"].data_ptr() if mom else 0, sh[0], p.numel(),
                         sh[1] | (uninit << 32)))
            key.append(rows[-1])
        key = tuple(key)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached
        tab = np.array(rows, dtype=np.int64).reshape(len(rows), 6)
        dev_tab = torch.from_numpy(tab).to(plist[0].device)
        total = int(sum(r[4] for r in rows))
        n_shadow = int(sum(r[4] for r in rows if r[3]))
        per_elem = 4 * (2 + 2 + (2 if mom else 0)) + 2.0 * n_shadow / max(total, 1)
        cached = (key, dev_tab, len(rows), total, per_elem)
        self._tables[gi] = cached
        self.table_builds += 1
        return cached

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        model = self._shadow_model()
        self._shadows_used = False
        touched = []
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda:
                    raise L.BvcError("FusedSGD runs on CUDA parameters only; there is no CPU path")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise L.BvcError("FusedSGD: fp32 dense parameters and gradients only")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise L.BvcError("FusedSGD: contiguous parameters and gradients only")
            _, tab, n, total, per_elem = self._table(gi, group, plist)
            gs = grad_scale.to(torch.float32).reshape(1) if grad_scale is not None else None
            fi = found_inf.to(torch.float32).reshape(1) if found_inf is not None else None
            with torch.cuda.device(plist[0].device):
                L.sgd_step(tab, n, total, float(group["lr"]), float(group["momentum"]), float(group["dampening"]),
                           float(group["weight_decay"]), bool(group["nesterov"]), gs, fi, per_elem)
            if group["momentum"] != 0:
                for p in plist:
                    st = self.state[p]
                    if st.get("_bvc_uninit", False) and found_inf is None:
                        st["_bvc_uninit"] = False
            touched.extend(plist)
        if found_inf is not None and any(self.state[p].get("_bvc_uninit", False) for p in touched):
            # a skipped first step leaves the buffers uninitialised; one host read, first step(s) only
            if float(found_inf) == 0.0:
                for p in touched:
                    self.state[p]["_bvc_uninit"] = False
        if touched:
            # the kernel wrote through raw pointers: tell autograd (saved-tensor checks) and the weight-copy cache
            torch.autograd.graph.increment_version(touched)
            if model is not None and self._shadows_used:
                model.weight_shadows_synced()  # a skipped step leaves both the weights and their copies untouched
        return loss

    def state_dict(self):
        sd = super().state_dict()
        sd["state"] = {k: {n: v for n, v in st.items() if n != "_bvc_uninit"} for k, st in sd["state"].items()}
        return sd
