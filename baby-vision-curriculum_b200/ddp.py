"""Data-parallel wrapper for the bvc VideoMAE model: torch DDP's role in the reference loop
(pretrain_videomae.py:180-181  `DDP(xmodel, device_ids=[rank], output_device=rank, find_unused_parameters=False)`),
re-designed around how this engine produces gradients.

torch's DistributedDataParallel (which also wraps the model unchanged, tests/test_model_gpu.py) copies every
parameter's gradient into its own flat buckets (one small kernel per parameter, ~200 per step), all-reduces the
buckets, and copies them back.  Here each backward stage (engine.py: EmbedFn / BlockFn / EncToDecFn / HeadLossFn)
already writes ALL of its parameter gradients into ONE contiguous fp32 buffer and hands autograd views of it, so that
buffer *is* the bucket: the stage launches one asynchronous NCCL all-reduce (average) on it the moment its kernels
are queued -- in reverse layer order, overlapped with the rest of the backward over NVLink -- and a callback at the
end of the backward makes the compute stream wait for the outstanding collectives.  No copy kernels, no extra
traffic, same result as DDP (mean of the per-rank gradients).

Same surface as the reference uses: `.module`, `.parameters()`, `forward(*a, **k)`, `no_sync()`; parameters are
broadcast from rank 0 at construction like DDP's _sync_module_states.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.distributed as dist
from torch import nn


class GradSync:
    """Per-model gradient synchroniser: stages call reduce(flat) from inside their backward."""

    def __init__(self, process_group=None, params=()):
        self.pg = process_group
        self.world = dist.get_world_size(process_group)
        self.enabled = True
        self.deferred = False           # set per forward: .grad tensors already exist (accumulation / set_to_none=False)
        self._params = list(params)
        self._works = []
        self._armed = False
        backend = dist.get_backend(process_group)
        self._avg = backend == "nccl"   # gloo (CPU tests) has no AVG: SUM then scale
        self.launched = 0               # collectives launched (bench / tests)
        # SMs left to NCCL while all-reduces are in flight (BVC_DDP_SM_RESERVE, default NCCL_MAX_CTAS if that is set):
        # the persistent kernels of the rest of the backward size their grids for the remaining SMs instead of
        # running their last CTAs, and those CTAs' share of the work, as a second wave (include/bvc.h bvc_set_sm_limit)
        self._sm_limit = 0
        if backend == "nccl" and torch.cuda.is_available():
            reserve = int(os.environ.get("BVC_DDP_SM_RESERVE", os.environ.get("NCCL_MAX_CTAS", "0")) or 0)
            n = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            if 0 < reserve < n // 2:
                self._sm_limit = n - reserve

    def reduce(self, flat: torch.Tensor):
        if not self.enabled or self.world == 1:
            return
        if not self._armed:
            # first bucket of this backward pass: have the engine call us when the pass is over
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
            self._armed = True
            if self._sm_limit:
                from . import _lib as L
                L.set_sm_limit(self._sm_limit)
        if self.deferred:
            return  # autograd will ADD this stage's gradients into existing .grad tensors: reduce those at the end
        self._launch(flat)

    def _launch(self, t):
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._works.append((dist.all_reduce(t, op=op, group=self.pg, async_op=True), t))
        self.launched += 1

    def begin_forward(self):
        """Overlapped mode needs autograd to adopt the stage buffers as the .grad tensors (zero_grad(set_to_none=True),
        the torch >= 2.0 default and what the reference's optimizer.zero_grad() does); if gradients are already
        allocated they are accumulated into, and only the accumulated result may be reduced."""
        self.deferred = any(p.grad is not None for p in self._params)

    def _finalize(self):
        if self.deferred:
            for p in self._params:
                if p.grad is not None:
                    self._launch(p.grad)
        works, self._works, self._armed = self._works, [], False
        if self._sm_limit:
            from . import _lib as L
            L.set_sm_limit(0)
        for w, flat in works:
            w.wait()  # stream-level on CUDA: the host does not block
            if not self._avg:
                flat.div_(self.world)


class DistributedDataParallel(nn.Module):
    def __init__(self, module, device_ids=None, output_device=None, dim=0, broadcast_buffers=True, process_group=None,
                 bucket_cap_mb=None, find_unused_parameters=False, check_reduction=False,
                 gradient_as_bucket_view=False, static_graph=False, **_ignored):
        super().__init__()
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("Default process group has not been initialized, please make sure to call "
                               "init_process_group.")  # torch DDP's message
        if find_unused_parameters:
            raise ValueError("every parameter of the VideoMAE pretraining step receives a gradient each step; "
                             "find_unused_parameters=True is not supported (the reference passes False)")
        if not hasattr(module, "_grad_sync"):
            raise TypeError("bvc_b200.DistributedDataParallel wraps bvc_b200.VideoMAEForPreTraining; use "
                            "torch.nn.parallel.DistributedDataParallel for other modules")
        self.module = module
        self.process_group = process_group
        self.device_ids = device_ids
        self.output_device = output_device
        self.sync = GradSync(process_group, module.parameters())
        module._grad_sync = self.sync
        # DDP semantics: every replica starts from rank 0's parameters (and buffers)
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)

    def forward(self, *args, **kwargs):
        if torch.is_grad_enabled():
            self.sync.begin_forward()
        return self.module(*args, **kwargs)

    @contextlib.contextmanager
    def no_sync(self):
        """Skip gradient synchronisation inside the context (gradient accumulation), like torch DDP's: the next
        synchronised backward finds the accumulated .grad tensors and reduces those (deferred mode)."""
        old = self.sync.enabled
        self.sync.enabled = False
        try:
            yield
        finally:
            self.sync.enabled = old
