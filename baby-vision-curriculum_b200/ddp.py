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
    """Per-model gradient synchroniser: stages call reduce(flat, params, views) from inside their backward.

    The mode of a backward pass is decided AT BACKWARD TIME, when the pass's first stage reports (nothing of this pass
    has reached a .grad yet): if no parameter has a .grad, autograd will adopt the stage views as the .grad tensors
    and every stage buffer is all-reduced in place, asynchronously, the moment it is complete ("overlapped"); if any
    .grad exists (gradient accumulation, zero_grad(set_to_none=False), a second backward through the same model before
    the first one's optimizer step), autograd ADDS the views into the existing tensors, so only the accumulated
    tensors may be reduced, at the end of the pass ("deferred" -- torch DDP's semantics after no_sync()).  Deciding at
    forward time was wrong for fwd, fwd, bwd, bwd: the second backward added local views into .grad tensors while
    their in-place all-reduce was still running.
    At the end of an overlapped pass every parameter's .grad is checked to alias the reduced view it was handed; a
    .grad that autograd copied instead of adopting (tensor hooks, create_graph) is overwritten with the reduced view."""

    def __init__(self, process_group=None, params=(), bucket_bytes=None):
        self.pg = process_group
        # Stage buffers are the buckets; `bucket_bytes` > 0 additionally COALESCES consecutive stages into one NCCL
        # group launch once that many bytes are pending (19 launches per ViT-B step become ~376.8 MB / bucket_bytes):
        # fewer NCCL kernels competing with the persistent GEMMs for SMs.  None: BVC_DDP_BUCKET_MB from the
        # environment, default 32 (ViT-B step on 8 x B200: 24.58 ms with 0, 24.21 with 32, 24.32 with 96, 24.72 with 400 --
        # inside the run-to-run noise of those boxes, profiles/r02_ddp_bucket_sweep.md; 7 launches instead of 19 is kept
        # for the fewer launches, not for a measured gain); 0: one collective per stage.
        if bucket_bytes is None:
            bucket_bytes = int(float(os.environ.get("BVC_DDP_BUCKET_MB", "32")) * (1 << 20))
        self.bucket_bytes = int(bucket_bytes)
        self._pending = []
        self._pending_bytes = 0
        self.world = dist.get_world_size(process_group)
        self.enabled = True
        self.deferred = False           # mode of the current / most recent backward pass
        self._params = list(params)
        self._works = []
        self._pairs = []                # (parameter, reduced view) of the overlapped stages of this pass
        self._armed = False
        backend = dist.get_backend(process_group)
        self._avg = backend == "nccl"   # gloo (CPU tests) has no AVG: SUM then scale
        self.launched = 0               # collectives launched (bench / tests)
        self.flushes = 0                # coalesced buckets flushed (one NCCL group launch each)
        self.adopted = 0                # .grad tensors found aliasing their reduced stage buffer (last pass)
        self.copied = 0                 # .grad tensors that had to be overwritten with the reduced view (last pass)

    def reduce(self, flat: torch.Tensor, params=None, views=None):
        if not self.enabled or self.world == 1:
            return
        if not self._armed:
            # first bucket of this backward pass: have the engine call us when the pass is over, and fix the mode
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
            self._armed = True
            self.deferred = any(p.grad is not None for p in self._params)
        if self.deferred:
            return  # autograd ADDS this stage's gradients into existing .grad tensors: reduce those at the end
        if self.bucket_bytes > 0:
            self._pending.append(flat)
            self._pending_bytes += flat.numel() * flat.element_size()
            if self._pending_bytes >= self.bucket_bytes:
                self._flush()
        else:
            self._launch(flat)
        if params is not None and views is not None:
            # NOT the view tensors themselves: a second reference keeps autograd's AccumulateGrad from adopting them
            # (it steals a gradient only when it holds the last reference) -- remember where they live instead
            for p, v in zip(params, views):
                if v is None:
                    continue
                self._pairs.append((p, v.data_ptr(), v.numel(), v.untyped_storage(), v.storage_offset()))

    def _launch(self, t):
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        self._works.append((dist.all_reduce(t, op=op, group=self.pg, async_op=True), t))
        self.launched += 1

    def _flush(self):
        """One NCCL group launch (ncclGroupStart / End around the pending stages' all-reduces) on NCCL; gloo has no
        coalescing for all_reduce: the same collectives one by one."""
        pend, self._pending, self._pending_bytes = self._pending, [], 0
        if not pend:
            return
        self.flushes += 1
        if len(pend) == 1 or not self._avg:
            for t in pend:
                self._launch(t)
            return
        op = dist.ReduceOp.AVG
        with dist._coalescing_manager(group=self.pg, device=pend[0].device, async_ops=True) as cm:
            for t in pend:
                dist.all_reduce(t, op=op, group=self.pg)
        self._works.append((cm, None))
        self.launched += 1

    def begin_forward(self):
        """Kept for callers of the round-1 interface: the mode is decided at backward time (see the class docstring)."""

    def _finalize(self):
        if self.deferred:
            for p in self._params:
                if p.grad is not None:
                    self._launch(p.grad)
        self._flush()
        works, self._works, self._armed = self._works, [], False
        pairs, self._pairs = self._pairs, []
        for w, flat in works:
            w.wait()  # stream-level on CUDA: the host does not block
            if not self._avg and flat is not None:
                flat.div_(self.world)
        adopted = copied = 0
        for p, ptr, n, storage, offset in pairs:
            g = p.grad
            if g is not None and g.data_ptr() == ptr and g.numel() == n:
                adopted += 1
            else:
                # autograd copied the view (on the compute stream, racing the in-place collective): the reduced view is
                # the gradient of this pass, and .grad was None when the pass began
                with torch.no_grad():
                    v = torch.empty(0, dtype=p.dtype, device=p.device).set_(storage, offset, p.shape)
                    if g is None:
                        p.grad = v.clone()
                    else:
                        g.copy_(v)
                copied += 1
        self.adopted, self.copied = adopted, copied


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
        self.sync = GradSync(process_group, module.parameters(),
                             None if bucket_cap_mb is None else int(bucket_cap_mb * (1 << 20)))
        module._grad_sync = self.sync
        # DDP semantics: every replica starts from rank 0's parameters (and buffers)
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t, src=dist.get_global_rank(process_group, 0) if process_group is not None else 0,
                               group=process_group)

    def forward(self, *args, **kwargs):
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
