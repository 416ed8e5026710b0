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


def test_allreduce_is_identity_without_process_group():
    sys.path.insert(0, ROOT)
    import bvc_b200 as bvc
    x = torch.tensor(2.0, requires_grad=True)
    y = bvc.AllReduce.apply(x * 1.0)
    y.backward()
    assert float(y) == 2.0 and float(x.grad) == 1.0


def test_reference_arm_prints_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
