"""World-size-2 gloo checks of the host-side distributed logic (no GPU): the AllReduce mirror of
pretraining/generative/ddputils.py:53-68 (forward = mean over ranks, backward = identity) and the reference arm's
"rank 0 alone prints" rule of bench.py."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bvc_b200 as bvc
    x = torch.tensor(float(rank + 1), requires_grad=True)
    y = bvc.AllReduce.apply(x * 1.0)
    (y * 3.0).backward()
    out[rank] = (float(y), float(x.grad))
    dist.destroy_process_group()


def test_allreduce_mirror_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29611, out), nprocs=2, join=True)
    # forward: (1 + 2) / 2 on both ranks; backward: identity (grad of the LOCAL loss, DDP averages separately)
    assert out[0] == (1.5, 3.0) and out[1] == (1.5, 3.0)


def _gather_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bvc_b200 as bvc
    sys.path.insert(0, "/root/reference/pretraining/predictive")
    x = (torch.arange(6, dtype=torch.float32).reshape(3, 2) + 10 * rank).requires_grad_(True)
    y = bvc.AllGather.apply(x)
    w = torch.arange(12, dtype=torch.float32).reshape(6, 2) * (rank + 1)     # a different upstream gradient per rank
    (y * w).sum().backward()
    res = {"y": y.detach().clone(), "gx": x.grad.clone()}
    try:  # the reference's own class, when the reference tree is present (build container)
        import distributed as RD
        xr = x.detach().clone().requires_grad_(True)
        yr = RD.AllGather.apply(xr)
        (yr * w).sum().backward()
        res["ref_equal"] = bool(torch.equal(yr, y) and torch.equal(xr.grad, x.grad))
    except ImportError:
        res["ref_equal"] = None
    out[rank] = res
    dist.destroy_process_group()


def test_allgather_mirror_world2():
    """bvc_b200.AllGather (predictive/distributed.py:49-76): forward = rows of every rank in rank order; backward =
    this rank's rows of the gradient summed over ranks -- and equal to the reference's class where that is importable."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_gather_worker, args=(2, 29613, out), nprocs=2, join=True)
    base = torch.arange(6, dtype=torch.float32).reshape(3, 2)
    want_y = torch.cat([base, base + 10])
    w_sum = torch.arange(12, dtype=torch.float32).reshape(6, 2) * 3          # (1 + 2) x the per-rank gradient
    for r in (0, 1):
        assert torch.equal(out[r]["y"], want_y)
        assert torch.equal(out[r]["gx"], w_sum[3 * r:3 * r + 3])
        assert out[r]["ref_equal"] in (True, None)


def test_allreduce_is_identity_without_process_group():
    sys.path.insert(0, ROOT)
    import bvc_b200 as bvc
    x = torch.tensor(2.0, requires_grad=True)
    y = bvc.AllReduce.apply(x * 1.0)
    y.backward()
    assert float(y) == 2.0 and float(x.grad) == 1.0
    z = torch.ones(2, 3, requires_grad=True)
    g = bvc.AllGather.apply(z)                       # identity (values and gradient) without a process group
    g.sum().backward()
    assert g.shape == (2, 3) and torch.equal(z.grad, torch.ones(2, 3))


def test_reference_arm_prints_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


# ------------------------------------------------------------------------------------------------------------------
# bvc_b200.DistributedDataParallel (ddp.py): stage gradient buffers are the all-reduce buckets.  The CUDA stages cannot
# run here, so a stand-in module produces its gradients the way engine.py does (views of one flat buffer, handed to
# GradSync.reduce from inside backward); the wrapper / GradSync logic under test is the shipped code.
class _StageFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, owner):
        ctx.owner, ctx.x = owner, x
        return (w * x).sum() + b.sum()

    @staticmethod
    def backward(ctx, g):
        flat = torch.zeros(6)
        gw, gb = flat[:4], flat[4:]
        gw.add_(ctx.x * g)
        gb.add_(g)
        if ctx.owner._grad_sync is not None:
            ctx.owner._grad_sync.reduce(flat, (ctx.owner.w, ctx.owner.b), (gw, gb))
        return None, gw, gb, None


class _StandIn(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(4))
        self.b = torch.nn.Parameter(torch.zeros(2))
        self._grad_sync = None

    def forward(self, x):
        return _StageFn.apply(x, self.w, self.b, self)


def _ddp_worker(rank, world, port, out, cap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bvc_b200 as bvc
    m = _StandIn()
    with torch.no_grad():
        m.w.fill_(float(rank + 5))  # replicas differ before wrapping: rank 0's values must win
    x = torch.arange(4.0) + 10 * rank
    ddp = bvc.DistributedDataParallel(m, device_ids=None, output_device=None, find_unused_parameters=False,
                                      bucket_cap_mb=cap)
    res = {"w_after_wrap": m.w.detach().clone().tolist(), "is_module": ddp.module is m,
           "n_params": len(list(ddp.parameters()))}
    # overlapped mode: .grad is None, autograd adopts the views of the stage buffer, which is all-reduced in place
    (ddp(x) * 2.0).backward()
    res["gw"], res["gb"] = m.w.grad.tolist(), m.b.grad.tolist()
    res["launched_overlap"] = ddp.sync.launched
    # deferred mode: gradients exist -> accumulate locally, reduce the accumulated tensors at the end of the pass
    (ddp(x) * 2.0).backward()
    res["gw2"] = m.w.grad.tolist()
    # no_sync: purely local accumulation
    m.zero_grad(set_to_none=True)
    with ddp.no_sync():
        ddp(x).backward()
    res["gw_local"] = m.w.grad.tolist()
    ddp(x).backward()  # accumulated (local + new) gets averaged, like torch DDP after no_sync
    res["gw_after_nosync"] = m.w.grad.tolist()
    # fwd, fwd, bwd, bwd: the mode is decided at BACKWARD time -- the first backward finds no .grad and
    # all-reduces its stage buffer in place (adopted as .grad), the second finds the .grad of the first and must
    # accumulate locally and reduce the accumulated tensors at the end (never add into a buffer that is being reduced)
    m.zero_grad(set_to_none=True)
    x2 = x + 100.0
    l1, l2 = ddp(x) * 2.0, ddp(x2) * 2.0
    l1.backward()
    res["adopted_1"], res["copied_1"], res["deferred_1"] = ddp.sync.adopted, ddp.sync.copied, ddp.sync.deferred
    l2.backward()
    res["deferred_2"] = ddp.sync.deferred
    res["gw_ffbb"], res["gb_ffbb"] = m.w.grad.tolist(), m.b.grad.tolist()
    # a tensor hook makes autograd hand AccumulateGrad a different tensor than the reduced view: the synchroniser
    # must notice that .grad does not alias its stage buffer and overwrite it with the reduced values
    m.zero_grad(set_to_none=True)
    h = m.w.register_hook(lambda g: g * 1.0)
    (ddp(x) * 2.0).backward()
    h.remove()
    res["gw_hook"], res["copied_hook"] = m.w.grad.tolist(), ddp.sync.copied
    out[rank] = res
    dist.destroy_process_group()


import pytest  # noqa: E402


@pytest.mark.parametrize("cap,port", [(None, 29612), (1.0, 29616)])
def test_bvc_ddp_world2_gloo(cap, port):
    """cap None: one collective per stage buffer; cap 1 MB: stage buffers are held back and flushed as one bucket
    (here at the end of the pass -- the stand-in's 24-byte stage never fills it) -- same gradients either way."""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, port, out, cap), nprocs=2, join=True)
    x0, x1 = torch.arange(4.0), torch.arange(4.0) + 10
    mean_gw = ((x0 + x1) / 2 * 2.0).tolist()
    for r in (0, 1):
        o = out[r]
        assert o["w_after_wrap"] == [5.0] * 4 and o["is_module"] and o["n_params"] == 2
        assert o["gw"] == mean_gw and o["gb"] == [2.0, 2.0]
        assert o["launched_overlap"] == 1          # one collective per stage buffer, not one per parameter
        # second backward accumulated into the averaged gradient: avg(g_avg + g_local) over ranks = 2 * g_avg
        assert o["gw2"] == [2 * v for v in mean_gw]
        assert o["gw_local"] == (x0 if r == 0 else x1).tolist()
        assert o["gw_after_nosync"] == ((x0 + x1) / 2 * 2).tolist()
        assert (o["adopted_1"], o["copied_1"], o["deferred_1"], o["deferred_2"]) == (2, 0, False, True)
        # both ranks: mean over ranks of (2 x) + mean over ranks of (2 (x + 100)) -- torch DDP's result for this order
        assert o["gw_ffbb"] == ((x0 + x1) / 2 * 2.0 + (x0 + x1 + 200) / 2 * 2.0).tolist() and o["gb_ffbb"] == [4.0, 4.0]
        assert o["gw_hook"] == mean_gw and o["copied_hook"] >= 1


def test_bvc_ddp_rejects_foreign_modules_and_unused_params():
    import pytest
    sys.path.insert(0, ROOT)
    import bvc_b200 as bvc
    with pytest.raises(RuntimeError):
        bvc.DistributedDataParallel(_StandIn())  # no process group


# ------------------------------------------------------------------------------------------------------------------
# coalesced buckets (GradSync.bucket_bytes): three stages whose buffers are 24 bytes each; with a 40-byte cap the first two
# stages are flushed together from inside backward, the third at the end of the pass -- same gradients as one collective
# per stage, fewer launches
class _ThreeStages(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.stages = torch.nn.ModuleList([_StandIn() for _ in range(3)])
        self._grad_sync = None

    def forward(self, x):
        for s in self.stages:
            s._grad_sync = self._grad_sync
        return sum((i + 1.0) * s(x) for i, s in enumerate(self.stages))


def _bucket_worker(rank, world, port, out, cap_bytes):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bvc_b200 as bvc
    m = _ThreeStages()
    ddp = bvc.DistributedDataParallel(m, bucket_cap_mb=cap_bytes / float(1 << 20))
    x = torch.arange(4.0) + 10 * rank
    ddp(x).backward()
    out[rank] = {"gw": [s.w.grad.tolist() for s in m.stages], "gb": [s.b.grad.tolist() for s in m.stages],
                 "launched": ddp.sync.launched, "flushes": ddp.sync.flushes, "adopted": ddp.sync.adopted,
                 "copied": ddp.sync.copied, "cap": ddp.sync.bucket_bytes}
    dist.destroy_process_group()


@pytest.mark.parametrize("cap_bytes,flushes,port", [(0, 0, 29621), (40, 2, 29622), (1 << 20, 1, 29623)])
def test_bvc_ddp_coalesced_buckets_world2(cap_bytes, flushes, port):
    """(gloo has no coalesced all_reduce: a flushed bucket is still one collective per stage buffer there, so the test
    counts bucket flushes -- on NCCL each flush is ONE group launch, tools/ddp_parity.py reports 4-7 per ViT-B step.)"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(2, port, out, cap_bytes), nprocs=2, join=True)
    x0, x1 = torch.arange(4.0), torch.arange(4.0) + 10
    for r in (0, 1):
        o = out[r]
        assert o["cap"] == cap_bytes and o["flushes"] == flushes and o["launched"] == 3
        assert (o["adopted"], o["copied"]) == (6, 0)
        for i in range(3):
            assert o["gw"][i] == ((i + 1.0) * (x0 + x1) / 2).tolist() and o["gb"][i] == [i + 1.0, i + 1.0]
