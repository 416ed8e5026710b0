#!/usr/bin/env python
"""Hardware parity of bvc_b200.DistributedDataParallel (ddp.py; the reference's
`DDP(xmodel, device_ids=[rank], output_device=rank, find_unused_parameters=False)`, pretrain_videomae.py:180-181) with
the REAL CUDA model on >= 2 GPUs over NCCL:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 \
        tools/ddp_parity.py [--config base --batch 8]

One step of the reference loop body (autocast + GradScaler-scaled backward) on rank-specific clips / masks, three ways:
  A  bvc_b200.DistributedDataParallel        (per-stage in-place all-reduce of the engine's gradient buffers)
  B  torch.nn.parallel.DistributedDataParallel around the same model class
  C  no wrapper: local gradients, then all_reduce(SUM) / world per parameter  (the definition of the DDP result)
Checks (one PASS / FAIL line each, exit code 1 on any FAIL):
  * A's gradients are BITWISE identical on every rank, every .grad aliases its reduced stage buffer (adopted == number
    of parameters, copied == 0);
  * A == C and B == C to rounding (the weight-gradient GEMMs accumulate split-K partials with fp32 atomics, so two runs
    of the same backward differ in the last bits: rel-L2 <= 2e-5 per tensor carrying >= 0.1 % of the norm, 1e-5 global);
  * fwd, fwd, bwd, bwd through A (gradient accumulation of two losses) == the sum of the two averaged gradients;
  * no_sync() accumulation followed by a synchronised backward == torch DDP's semantics;
  * several optimizer steps (GradScaler + FusedSGD with the bf16 weight copies updated in the same pass): parameters
    stay bitwise identical across ranks and the loss trajectory equals torch DDP + torch.optim.SGD's to 1e-4.
Also (BASELINE.json config 3): NT-Xent over embeddings gathered across the GPUs -- bvc_b200.AllGather
(predictive/distributed.py:49-76) + bvc_b200.info_nce_loss -- against the fp64 oracle on the concatenated features.
"""
import argparse
import copy
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# The A / B / C comparisons run SEPARATE backward passes and expect them to agree to fp32 rounding, so this script selects
# the deterministic two-pass attention backward (the default one-pass kernel reduces dQ with fp32 L2 atomics, whose
# order varies run to run: ~2e-5 global, like torch's flash backward).  The wrapper logic under test does not depend on
# which attention kernel produced the gradients.  (Read once by libbvc.so at its first attention call.)
os.environ.setdefault("BVC_ATTN_BWD1", "0")

FAILS = []


def report(name, ok, detail=""):
    if dist.get_rank() == 0:
        print(("PASS " if ok else "FAIL ") + name + ("  " + detail if detail else ""), flush=True)
    if not ok:
        FAILS.append(name)


def all_ranks(flag: bool) -> bool:
    t = torch.tensor([1.0 if flag else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item() > 0)


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / max(float(b.double().norm()), 1e-30))


def compare(ga, gc, tag, per_tensor=2e-5, glob=1e-5):
    num = den = 0.0
    worst = ("", 0.0)
    tot = sum(float(v.double().pow(2).sum()) for v in gc.values()) ** 0.5
    for k, r in gc.items():
        e = float((ga[k].double() - r.double()).norm())
        n = float(r.double().norm())
        num += e * e
        den += n * n
        if n >= 1e-3 * tot and e / max(n, 1e-30) > worst[1]:
            worst = (k, e / max(n, 1e-30))
    g = (num / max(den, 1e-60)) ** 0.5
    report(tag, all_ranks(g <= glob and worst[1] <= per_tensor), f"global rel-L2 {g:.2e}, worst tensor {worst[0]} {worst[1]:.2e}")


def bitwise_equal_across_ranks(tensors, tag):
    ok = True
    for t in tensors:
        ref = t.detach().clone()
        dist.broadcast(ref, src=0)
        ok = ok and torch.equal(ref, t.detach())
    report(tag, all_ranks(ok))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="base")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import bvc_b200 as bvc
    from bench import CONFIGS, GRID, MASK_RATIO

    cfg = bvc.VideoMAEConfig(**CONFIGS[a.config])
    torch.manual_seed(1000 + rank)  # replicas DIFFER before wrapping: rank 0's values must win (DDP broadcast)
    base = bvc.VideoMAEForPreTraining(cfg).to(dev).train()
    g = torch.Generator().manual_seed(77 + rank)
    clips = [torch.randn(a.batch, 16, 3, 224, 224, generator=g).to(dev) for _ in range(2)]
    np.random.seed(500 + rank)
    masks = [bvc.batch_masks(bvc.TubeMaskingGenerator(GRID, MASK_RATIO), a.batch).to(dev) for _ in range(2)]
    n_params = len(list(base.parameters()))

    def grads_of(model):
        return {k: p.grad.detach().clone() for k, p in model.named_parameters()}

    def fwd_bwd(m, i, scale=1024.0):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = m(clips[i], bool_masked_pos=masks[i]).loss
        (loss * scale).backward()
        return loss.detach()

    # ---------------------------------------------------------------- A: bvc DDP
    model_a = copy.deepcopy(base)
    ddp_a = bvc.DistributedDataParallel(model_a, device_ids=[local], output_device=local, find_unused_parameters=False)
    bitwise_equal_across_ranks(list(model_a.parameters()), "A: parameters broadcast from rank 0 at construction")
    state0 = copy.deepcopy(model_a.state_dict())
    fwd_bwd(ddp_a, 0)
    torch.cuda.synchronize()
    ga = grads_of(model_a)
    report("A: every .grad aliases its all-reduced stage buffer (adopted == n_params, copied == 0)",
           all_ranks(ddp_a.sync.adopted == n_params and ddp_a.sync.copied == 0 and not ddp_a.sync.deferred),
           f"adopted {ddp_a.sync.adopted} / {n_params}, copied {ddp_a.sync.copied}, collectives {ddp_a.sync.launched}")
    bitwise_equal_across_ranks(list(ga.values()), "A: gradients bitwise identical on every rank")

    # ---------------------------------------------------------------- C: the definition
    model_c = copy.deepcopy(base)
    model_c.load_state_dict(state0)
    fwd_bwd(model_c, 0)
    gc = {}
    for k, p in model_c.named_parameters():
        t = p.grad.detach().clone()
        dist.all_reduce(t)
        gc[k] = t / world
    compare(ga, gc, "A (bvc DDP) == C (mean of the per-rank gradients)")

    # ---------------------------------------------------------------- B: torch DDP
    model_b = copy.deepcopy(base)
    model_b.load_state_dict(state0)
    ddp_b = torch.nn.parallel.DistributedDataParallel(model_b, device_ids=[local], output_device=local,
                                                      find_unused_parameters=False)
    fwd_bwd(ddp_b, 0)
    torch.cuda.synchronize()
    gb = grads_of(model_b)
    compare(gb, gc, "B (torch DDP around the bvc model) == C")
    compare(ga, gb, "A == B")

    # ---------------------------------------------------------------- fwd, fwd, bwd, bwd through A
    model_a.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        l0 = ddp_a(clips[0], bool_masked_pos=masks[0]).loss
        l1 = ddp_a(clips[1], bool_masked_pos=masks[1]).loss
    (l0 * 1024.0).backward()
    first_mode = ddp_a.sync.deferred
    (l1 * 1024.0).backward()
    torch.cuda.synchronize()
    g_ffbb = grads_of(model_a)
    model_c.zero_grad(set_to_none=True)
    fwd_bwd(model_c, 0)
    fwd_bwd(model_c, 1)
    gc2 = {}
    for k, p in model_c.named_parameters():
        t = p.grad.detach().clone()
        dist.all_reduce(t)
        gc2[k] = t / world
    report("A: fwd, fwd, bwd, bwd -- first backward overlapped, second deferred",
           all_ranks(first_mode is False and ddp_a.sync.deferred is True))
    compare(g_ffbb, gc2, "A: fwd, fwd, bwd, bwd == mean of the accumulated per-rank gradients")
    bitwise_equal_across_ranks(list(g_ffbb.values()), "A: accumulated gradients bitwise identical on every rank")

    # ---------------------------------------------------------------- no_sync accumulation
    model_a.zero_grad(set_to_none=True)
    with ddp_a.no_sync():
        fwd_bwd(ddp_a, 0)
    fwd_bwd(ddp_a, 1)
    torch.cuda.synchronize()
    compare(grads_of(model_a), gc2, "A: no_sync() accumulation + synchronised backward == C")

    # ---------------------------------------------------------------- training steps: A + FusedSGD vs B + torch SGD
    for m in (model_a, model_b):
        m.load_state_dict(state0)
        m.zero_grad(set_to_none=True)
    opt_a = bvc.FusedSGD(ddp_a.parameters(), lr=0.1, momentum=0.9, nesterov=True, shadow_from=model_a)
    opt_b = torch.optim.SGD(ddp_b.parameters(), lr=0.1, momentum=0.9, nesterov=True)
    sc_a, sc_b = torch.amp.GradScaler("cuda"), torch.amp.GradScaler("cuda")
    la, lb = [], []
    for s in range(a.steps):
        for xm, opt, sc, out in ((ddp_a, opt_a, sc_a, la), (ddp_b, opt_b, sc_b, lb)):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                opt.zero_grad()
                loss = bvc.AllReduce.apply(xm(clips[s % 2], bool_masked_pos=masks[s % 2]).loss)
            sc.scale(loss).backward()
            sc.step(opt)
            sc.update()
            out.append(float(loss))
    torch.cuda.synchronize()
    bitwise_equal_across_ranks(list(model_a.parameters()), f"A: parameters bitwise identical on every rank after {a.steps} steps")
    chk = torch.stack([p.detach().double().sum() for p in model_a.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    report("A: parameter checksum equal across ranks", float(lo) == float(hi), f"{float(chk):.9e}")
    dl = max(abs(x - y) / abs(y) for x, y in zip(la, lb))
    report("A + FusedSGD loss trajectory == B + torch.optim.SGD", all_ranks(dl <= 1e-4),
           f"max rel diff {dl:.2e}; A {['%.6f' % v for v in la]} B {['%.6f' % v for v in lb]}")
    compare({k: p.detach() for k, p in model_a.named_parameters()}, {k: p.detach() for k, p in model_b.named_parameters()},
            f"parameters after {a.steps} steps: A == B", per_tensor=1e-4, glob=1e-5)

    # ---------------------------------------------------------------- the same steps through ONE captured CUDA graph
    # bvc_b200.GraphedTrainStep under bvc DDP: the stage all-reduces are captured with the kernels; the replayed loop must
    # reproduce the eager loop above (same start state, same batches) and keep the ranks bitwise in step
    del l0, l1  # loss tensors of the eager sections: their autograd graphs pin model_a's AccumulateGrad nodes to the
    #             legacy default stream, which no capture can include (see graphed.py)
    model_a.load_state_dict(state0)
    model_a.zero_grad(set_to_none=True)
    opt_g = bvc.FusedSGD(ddp_a.parameters(), lr=0.1, momentum=0.9, nesterov=True, shadow_from=model_a)
    gstep = bvc.GraphedTrainStep(ddp_a, opt_g, torch.amp.GradScaler("cuda"), loss_fn=bvc.AllReduce.apply, warmup=2)
    lg = []
    for s in range(a.steps + 2):
        lg.append(float(gstep(clips[s % 2], masks[s % 2])))
    torch.cuda.synchronize()
    report(f"A + GraphedTrainStep: {gstep.replays} of {a.steps + 2} steps replayed from one capture",
           all_ranks(gstep.captures == 1 and gstep.replays == a.steps), f"captures {gstep.captures}")
    bitwise_equal_across_ranks(list(model_a.parameters()),
                               f"A + GraphedTrainStep: parameters bitwise identical on every rank after {a.steps + 2} steps")
    dlg = max(abs(x - y) / abs(y) for x, y in zip(lg[:a.steps], la))
    report("A + GraphedTrainStep loss trajectory == the eager loop's", all_ranks(dlg <= 1e-4),
           f"max rel diff {dlg:.2e}; graphed {['%.6f' % v for v in lg]}")
    # a captured graph holds NCCL work of the communicator: release it before the process group is torn down
    del gstep, opt_g
    import gc as _gc
    _gc.collect()
    torch.cuda.synchronize()

    # ---------------------------------------------------------------- config 3: NT-Xent over gathered embeddings
    from oracle import simclr_oracle as SO
    nloc, D, T = 256, 512, 0.1
    gf = torch.Generator().manual_seed(9 + rank)
    f = torch.randn(nloc, D, generator=gf)
    f[1::2] = 0.7 * f[0::2] + 0.3 * f[1::2]
    floc = f.to(dev).requires_grad_(True)
    feats = bvc.AllGather.apply(floc)
    n = feats.shape[0]
    masks_nce = bvc.make_simclr_masks(n, dev) if hasattr(bvc, "make_simclr_masks") else None
    loss = bvc.info_nce_loss(T, masks_nce, feats)
    (loss * (rank + 1.0)).backward()   # a different upstream gradient per rank, like per-rank losses
    torch.cuda.synchronize()
    gathered = [torch.zeros(nloc, D, device=dev) for _ in range(world)]
    dist.all_gather(gathered, floc.detach())
    fall = torch.cat(gathered).cpu().double().requires_grad_(True)
    pos, neg = SO.make_masks(n)
    ref = SO.info_nce_loss(T, (pos, neg), fall)
    w = sum(r + 1.0 for r in range(world))   # AllGather.backward sums the ranks' upstream gradients
    (ref * w).backward()
    gref = fall.grad[rank * nloc:(rank + 1) * nloc]
    e_loss = abs(float(loss) - float(ref)) / abs(float(ref))
    e_grad = rel_l2(floc.grad.cpu(), gref)
    report(f"config 3: AllGather + info_nce_loss over {world} x {nloc} embeddings vs fp64 oracle",
           all_ranks(e_loss <= 1e-4 and e_grad <= 1e-3), f"loss rel {e_loss:.2e}, local-gradient rel-L2 {e_grad:.2e}")

    if rank == 0:
        print(f"DDP-PARITY world {world}: {'ALL PASS' if not FAILS else 'FAILED: ' + '; '.join(FAILS)}", flush=True)
    dist.barrier()
    torch.cuda.synchronize()
    code = 1 if FAILS else 0
    # communicator teardown has been seen to hang after CUDA-graph captures of NCCL work (600 s of a 2-GPU box): the
    # verdict is printed, every rank is past the barrier -- leave without it if it does not return promptly
    import threading
    sys.stdout.flush()
    sys.stderr.flush()
    threading.Timer(20.0, lambda: os._exit(code)).start()
    dist.destroy_process_group()
    os._exit(code)


if __name__ == "__main__":
    main()
