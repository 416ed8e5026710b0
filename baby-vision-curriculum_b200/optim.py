"""Fused optimizers on libbvc.so multi-tensor kernels (csrc/optim.cu) with torch.optim's interfaces:

  FusedSGD    torch.optim.SGD   -- pretrain_videomae.py:187-189 (nesterov, momentum 0.9, lr 0.1, wd 0)     bvc_sgd_step
  FusedAdamW  torch.optim.AdamW -- pretrain_videomae.py:190-191 (betas (0.9, 0.95))                       bvc_adam_step
  FusedAdam   torch.optim.Adam  -- pretrain_videomae.py:192-193                                           bvc_adam_step

Drop-ins for the reference's optimizer lines; state-dict layouts are torch's (`momentum_buffer`; `step`, `exp_avg`,
`exp_avg_sq` per parameter).  Under torch.amp.GradScaler (pretrain_videomae.py:312-314) the whole of
`scaler.step(optimizer)` is two launches: GradScaler hands itself over through the `grad_scaler` argument of step()
(torch/amp/grad_scaler.py: "their step() should accept an additional, optional grad_scaler kwarg"), the inf / nan check
is ONE read-only multi-tensor kernel (bvc_grad_nonfinite; torch's _amp_foreach_non_finite_check_and_unscale_ takes 5
launches and re-writes every gradient), and the update kernel unscales (writing the unscaled gradient back, which is
what the reference's grad logger reads after scaler.step -- loggingtools.py:107-118), skips on overflow, updates, and
refreshes the model's bf16 operand copies (`shadow_from=model`), all on the device.  If a future torch stops passing
`grad_scaler`, the `grad_scale` / `found_inf` attribute protocol (`_step_supports_amp_scaling`) is honoured as well.
No CPU path: parameters must live on a CUDA device.
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from . import _lib as L

# GradScaler announces (once per call site) that it will stop passing itself; both protocols are implemented here
warnings.filterwarnings("ignore", message="GradScaler is going to stop passing itself", category=FutureWarning)


class _FusedBase(torch.optim.Optimizer):
    _step_supports_amp_scaling = True
    _state_keys = ()      # per-parameter state tensors shaped like the parameter; the first one is the table's `m`

    def __init__(self, params, defaults, shadow_from=None):
        super().__init__(params, defaults)
        self._shadow_from = shadow_from
        self._tables = {}        # group index -> (key, device tables, n_entries, total elements, shadowed elements)
        self._pinned = {}        # group index -> pinned host staging of the pointer table
        self._pin_events = {}    # group index -> event after the last copy out of the staging buffer
        self.table_builds = 0    # how often a pointer table had to be rebuilt (stable pointers -> stays small)
        self._found_inf = {}     # device -> fp32 scalar written by bvc_grad_nonfinite
        self._graph_pins = []    # pinned pointer tables referenced by captured CUDA graphs

    def _shadow_model(self):
        m = self._shadow_from
        if m is None:
            return None
        return getattr(m, "module", m)

    def _needs_state(self, group):
        return True

    def _ensure_state(self, p, group):
        """Allocate missing state; returns 1 while the buffers hold nothing yet (first applied step initialises them)."""
        st = self.state[p]
        if not self._needs_state(group):
            return 0
        k0 = self._state_keys[0]
        if st.get(k0) is None:  # absent, or torch's `momentum_buffer: None` of a state dict saved before any step
            for k in self._state_keys:
                st[k] = torch.empty_like(p, memory_format=torch.preserve_format)
            st["_bvc_uninit"] = True
        return 1 if st.get("_bvc_uninit", False) else 0

    def _table(self, gi, group, plist):
        model = self._shadow_model()
        shadows = model.weight_shadows() if model is not None else {}  # {} unless the copies match the weights now
        if shadows:
            self._shadows_used = True
        has_state = self._needs_state(group)
        rows, extra = [], []
        for p in plist:
            uninit = self._ensure_state(p, group)
            st = self.state[p]
            sh = shadows.get(id(p), (0, 0))
            rows.append((p.data_ptr(), p.grad.data_ptr(), st[self._state_keys[0]].data_ptr() if has_state else 0, sh[0],
                         p.numel(), sh[1] | (uninit << 32)))
            extra.append(tuple(st[k].data_ptr() for k in self._state_keys[1:]) if has_state else ())
        key = (tuple(rows), tuple(extra))
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached
        # pointer tables staged in pinned memory and copied without blocking: the caching allocator hands the stage
        # gradient buffers new addresses now and then, and a pageable copy here was a hidden synchronisation
        n, n_extra = len(rows), (len(self._state_keys) - 1 if has_state else 0)
        words = 6 * n + n_extra * n
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            # CUDA-graph capture: the copy below becomes a node that re-reads its pinned source at every replay, and no
            # event may be queried or synchronised -- a staging buffer of its own, kept alive with the optimizer
            pin = torch.empty(max(words, 64), dtype=torch.int64).pin_memory()
            self._graph_pins.append(pin)
        else:
            ev = self._pin_events.get(gi)
            if ev is not None:
                ev.synchronize()  # the previous copy out of the staging buffer (long done)
            pin = self._pinned.get(gi)
            if pin is None or pin.numel() < words:
                pin = self._pinned[gi] = torch.empty(max(words, 64), dtype=torch.int64).pin_memory()
        host = pin.numpy()
        host[:6 * n] = np.array(rows, dtype=np.int64).reshape(-1)
        for j in range(n_extra):
            host[6 * n + j * n:6 * n + (j + 1) * n] = [e[j] for e in extra]
        dev = plist[0].device
        dev_tab = torch.empty(words, dtype=torch.int64, device=dev)
        dev_tab.copy_(pin[:words], non_blocking=True)
        if not capturing:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            self._pin_events[gi] = ev
        total = int(sum(r[4] for r in rows))
        n_shadow = int(sum(r[4] for r in rows if r[3]))
        cached = (key, dev_tab, n, total, n_shadow)
        self._tables[gi] = cached
        self.table_builds += 1
        return cached

    def _check_params(self, plist):
        name = type(self).__name__
        for p in plist:
            if not p.is_cuda:
                raise L.BvcError(f"{name} runs on CUDA parameters only; there is no CPU path")
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                raise L.BvcError(f"{name}: fp32 dense parameters and gradients only")
            if not p.is_contiguous() or not p.grad.is_contiguous():
                raise L.BvcError(f"{name}: contiguous parameters and gradients only")

    def _launch(self, group, tab, n, total, n_shadow, gs, fi):  # pragma: no cover - subclasses
        raise NotImplementedError

    @torch.no_grad()
    def step(self, closure=None, grad_scaler=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        model = self._shadow_model()
        self._shadows_used = False
        work = []
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            self._check_params(plist)
            work.append((gi, group, plist, self._table(gi, group, plist)))
        if grad_scaler is not None and work:
            grad_scale, found_inf = self._scaler_handshake(grad_scaler, work)
        touched = []
        for gi, group, plist, (_, tab, n, total, n_shadow) in work:
            gs = grad_scale.to(torch.float32).reshape(1) if grad_scale is not None else None
            fi = found_inf.to(torch.float32).reshape(1) if found_inf is not None else None
            with torch.cuda.device(plist[0].device):
                self._launch(group, tab, n, total, n_shadow, gs, fi)
            if self._needs_state(group) and found_inf is None:
                for p in plist:
                    self.state[p]["_bvc_uninit"] = False
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

    def _scaler_handshake(self, scaler, work):
        """GradScaler.step passed itself (`grad_scaler`): do what its READY-stage branch would have done --
        _check_inf_per_device + scale / found_inf -- with one read-only kernel per device, and record found_inf where
        GradScaler.update() looks for it."""
        from torch.amp.grad_scaler import OptState
        state = scaler._per_optimizer_states[id(self)]
        if state["stage"] is OptState.UNSCALED:   # the user called scaler.unscale_(optimizer): already checked
            fi = sum(t.to(torch.float32) for t in state["found_inf_per_device"].values())
            return None, fi
        scale = scaler._get_scale_async()
        if scale is None:  # lazily initialised by scaler.scale(); step() without scale() is the user's error
            raise RuntimeError("GradScaler.step(optimizer) called before GradScaler.scale(loss)")
        per_dev = {}
        for gi, group, plist, (_, tab, n, total, _) in work:
            dev = plist[0].device
            fi = self._found_inf.get(dev)
            if fi is None:
                fi = self._found_inf[dev] = torch.zeros((), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                if dev in per_dev:  # several groups on one device: accumulate
                    tmp = torch.zeros((), dtype=torch.float32, device=dev)
                    L.grad_nonfinite(tab, n, total, tmp)
                    fi.add_(tmp)
                else:
                    L.grad_nonfinite(tab, n, total, fi)
            per_dev[dev] = fi
        state["found_inf_per_device"] = per_dev
        fi = per_dev[scale.device] if len(per_dev) == 1 and scale.device in per_dev else \
            sum(t.to(scale.device, non_blocking=True) for t in per_dev.values())
        return scale, fi

    def state_dict(self):
        sd = super().state_dict()
        sd["state"] = {k: {n: v for n, v in st.items() if n != "_bvc_uninit"} for k, st in sd["state"].items()}
        return sd


class FusedSGD(_FusedBase):
    _state_keys = ("momentum_buffer",)

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
                                      nesterov=nesterov), shadow_from)

    def _needs_state(self, group):
        return group["momentum"] != 0

    def _launch(self, group, tab, n, total, n_shadow, gs, fi):
        per_elem = 4 * (2 + 2 + (2 if group["momentum"] != 0 else 0)) + 2.0 * n_shadow / max(total, 1)
        L.sgd_step(tab, n, total, float(group["lr"]), float(group["momentum"]), float(group["dampening"]),
                   float(group["weight_decay"]), bool(group["nesterov"]), gs, fi, per_elem)


class FusedAdam(_FusedBase):
    """torch.optim.Adam's interface (L2 weight decay folded into the gradient)."""
    _state_keys = ("exp_avg", "exp_avg_sq")
    _decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, *,
                 maximize=False, shadow_from=None):
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")  # torch/optim/adam.py
        if amsgrad or maximize:
            raise NotImplementedError("amsgrad / maximize are not on the reference's path")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay), shadow_from)
        self._steps = {}  # group index -> device fp32 step count, shared by the group's parameters as state['step']

    def _step_tensor(self, gi, group, plist):
        t = self._steps.get(gi)
        if t is None:
            # after load_state_dict the parameters carry their own `step` tensors: they agree within a group
            loaded = [self.state[p]["step"] for p in plist if torch.is_tensor(self.state[p].get("step"))]
            v = float(loaded[0]) if loaded else 0.0
            t = self._steps[gi] = torch.full((), v, dtype=torch.float32, device=plist[0].device)
        for p in plist:
            self.state[p]["step"] = t
        return t

    def _launch(self, group, tab, n, total, n_shadow, gs, fi):
        gi = next(i for i, g in enumerate(self.param_groups) if g is group)
        plist = [p for p in group["params"] if p.grad is not None]
        step = self._step_tensor(gi, group, plist)
        per_elem = 4 * 8 + 2.0 * n_shadow / max(total, 1)
        L.adam_step(tab[:6 * n], tab[6 * n:7 * n], n, total, float(group["lr"]), float(group["betas"][0]),
                    float(group["betas"][1]), float(group["eps"]), float(group["weight_decay"]), self._decoupled, step,
                    gs, fi, per_elem)

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._steps = {}
        self._tables = {}


class FusedAdamW(FusedAdam):
    """torch.optim.AdamW's interface (decoupled weight decay, default 1e-2)."""
    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 maximize=False, shadow_from=None):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad,
                         maximize=maximize, shadow_from=shadow_from)
