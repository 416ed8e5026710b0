"""The reference loop body (pretrain_videomae.py:292-314: zero_grad, forward under autocast, scalar-loss all-reduce,
GradScaler-scaled backward, scaler.step, scaler.update) captured into ONE CUDA graph and replayed.

Why: a step is ~360 kernel launches issued from Python through ctypes -- 11.7 ms of host time per step on the GPU box,
whatever the batch.  At the BASELINE batch (64 clips, 23 ms of GPU work) the host stays ahead; at the batch the
reference's own SLURM scripts use (16 per GPU, slurm_dev_def.bash) the GPU needs 7.5 ms and the loop is launch-bound:
measured on B200 11.7 ms per step eager -> 7.5 ms replayed (4 clips: 11.9 -> 4.3 ms).  The replayed step also does not
care about host stalls (the multi-GPU passes that a 100 ms stall on one rank slows down for every rank).

    step = bvc.GraphedTrainStep(xmodel, optimizer, scaler, loss_fn=bvc.AllReduce.apply)
    for inputs, mask in loader:                      # fixed shapes (the reference's DataLoader has drop_last=True)
        loss = step(inputs, mask)                    # same arithmetic, same order as the eager loop body

The first `warmup` calls run the eager loop body (they are ordinary training steps: lazy initialisation, optimizer
state, the one-time read-back of the visible-token count); the next call copies its inputs into static buffers,
captures and replays; every later call is two device copies + one graph launch.  `loss` is a static tensor that the next
call overwrites.  Parameter .grad tensors live in the graph's memory pool and hold the step's unscaled gradients
after every call, as after the eager `scaler.step` (what the reference's grad logger reads, loggingtools.py:107-118).

What a captured graph freezes, and how this class deals with it:
  * shapes / dtypes / device of the inputs -- a call with other shapes runs eagerly (and says so once);
  * optimizer hyper-parameters (lr, momentum, betas, weight decay are kernel arguments) -- compared on every call, a
    change (LR schedule) triggers a re-capture;
  * the visible-token count per row -- `static_mask_count` is switched on: validated on the device every step, a
    mismatch poisons that step's loss with NaN and is raised by `model.check_mask_status()`.
Works under bvc_b200.DistributedDataParallel (the NCCL all-reduces of the stage buffers are captured with the rest).

Streams: the eager calls and the capture run on a private side stream (ordered after / before the caller's stream), as
the torch.cuda.graphs recipe asks -- autograd binds every parameter's AccumulateGrad node to the stream it was created
on, and a node bound to the legacy default stream cannot take part in a capture.  Those nodes are re-created by every
forward UNLESS an older autograd graph is still alive: drop the loss tensors of earlier eager iterations (`del loss`)
before the first graphed call, or the capture fails with cudaErrorStreamCaptureImplicit.
"""
from __future__ import annotations

import warnings

import torch

from . import _lib as L


class GraphedTrainStep:
    def __init__(self, model, optimizer, scaler=None, loss_fn=None, warmup=3, autocast_dtype=torch.bfloat16):
        if warmup < 2:
            raise ValueError("warmup >= 2: the first two optimizer steps initialise state on the host side")
        self.model, self.optimizer, self.scaler, self.loss_fn = model, optimizer, scaler, loss_fn
        self.warmup, self.dtype = int(warmup), autocast_dtype
        self.calls = 0            # steps taken through this object (eager + replayed)
        self.captures = 0
        self.replays = 0
        self._graph = None
        self._key = None          # (input signature, optimizer hyper-parameters) of the captured graph
        self._x = self._m = self._loss = None
        self._grads = None
        self._warned = False
        self._stale_grads = False
        self._stream = None       # private side stream of the eager calls and the capture
        base = getattr(model, "module", model)
        if not hasattr(base, "static_mask_count"):
            raise TypeError("GraphedTrainStep drives bvc_b200.VideoMAEForPreTraining (optionally DDP-wrapped)")
        base.static_mask_count = True
        self._base = base

    # the reference's loop body, verbatim (pretrain_videomae.py:292-314)
    def _eager(self, x, m):
        with torch.autocast("cuda", dtype=self.dtype):
            self.optimizer.zero_grad()
            loss = self.model(x, bool_masked_pos=m).loss
            if self.loss_fn is not None:
                loss = self.loss_fn(loss)
        if self.scaler is not None:
            self.scaler.scale(loss).backward()
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            loss.backward()
            self.optimizer.step()
        # detached: the step is complete, and a loss kept alive with its autograd graph would pin this iteration's
        # AccumulateGrad nodes (and their stream) into the next one
        return loss.detach()

    def _on_side_stream(self, fn, dev):
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        self._stream.wait_stream(cur)
        with torch.cuda.stream(self._stream):
            out = fn()
        cur.wait_stream(self._stream)
        return out

    def _signature(self, x, m):
        hp = tuple(tuple(sorted((k, repr(v)) for k, v in g.items() if k != "params")) for g in self.optimizer.param_groups)
        return (tuple(x.shape), x.dtype, tuple(m.shape), m.dtype, str(x.device), hp)

    def _capture(self, x, m):
        self._x = torch.empty_like(x)
        self._m = torch.empty_like(m)
        self._x.copy_(x)
        self._m.copy_(m)
        torch.cuda.synchronize(x.device)
        graph = torch.cuda.CUDAGraph()
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=x.device)
        # thread_local: the capture must not trip over harmless calls of other threads (DataLoader pin-memory thread)
        try:
            with torch.cuda.graph(graph, stream=self._stream, capture_error_mode="thread_local"):
                self._loss = self._eager(self._x, self._m)
        except Exception as e:
            raise L.BvcError("GraphedTrainStep: the training step could not be captured into a CUDA graph. If the error "
                             "is cudaErrorStreamCaptureImplicit, an autograd graph of an earlier eager iteration is still "
                             "alive (a loss tensor kept in a variable): delete it before the first graphed call. "
                             f"[{type(e).__name__}: {str(e).splitlines()[0]}]") from e
        self._graph = graph
        self._grads = [(p, p.grad) for g in self.optimizer.param_groups for p in g["params"]]
        self.captures += 1

    def __call__(self, pixel_values, bool_masked_pos):
        if not (pixel_values.is_cuda and bool_masked_pos.is_cuda):
            raise L.BvcError("GraphedTrainStep: inputs and masks must be CUDA tensors (there is no CPU path; copy them "
                             "with .to(device, non_blocking=True) as the reference loop does)")
        self.calls += 1
        if self.calls <= self.warmup:
            return self._on_side_stream(lambda: self._eager(pixel_values, bool_masked_pos), pixel_values.device)
        key = self._signature(pixel_values, bool_masked_pos)
        if self._graph is not None and key != self._key:
            same_inputs = key[:5] == self._key[:5]
            if same_inputs:          # optimizer hyper-parameters changed: freeze the new values
                self._graph = None
            else:
                if not self._warned:
                    warnings.warn("GraphedTrainStep: input shapes differ from the captured ones; this call runs eagerly")
                    self._warned = True
                out = self._on_side_stream(lambda: self._eager(pixel_values, bool_masked_pos), pixel_values.device)
                self._stale_grads = True
                return out
        if self._graph is None:
            self._capture(pixel_values, bool_masked_pos)
            self._key = key
        else:
            self._x.copy_(pixel_values, non_blocking=True)
            self._m.copy_(bool_masked_pos, non_blocking=True)
        self._graph.replay()
        self.replays += 1
        if self._stale_grads:   # an eager call in between re-pointed .grad at its own tensors
            for p, g in self._grads:
                p.grad = g
            self._stale_grads = False
        return self._loss
