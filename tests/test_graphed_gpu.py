"""bvc_b200.GraphedTrainStep: the reference loop body (pretrain_videomae.py:292-314) captured into one CUDA graph must
take the SAME training steps as the eager loop -- loss trajectory, gradients left in .grad, parameters -- including
across an optimizer hyper-parameter change (re-capture) and a call with other input shapes (eager fallback)."""
import warnings

import numpy as np
import pytest
import torch

from oracle import videomae_oracle as O
from tests.helpers import bvc_config, rel_l2

pytestmark = pytest.mark.gpu


def _setup(cfg, params, opt_name):
    import bvc_b200 as bvc
    m = bvc.VideoMAEForPreTraining(bvc_config(cfg))
    m.load_state_dict(params, strict=True)
    m = m.to("cuda:0").train()
    if opt_name == "sgd":
        o = bvc.FusedSGD(m.parameters(), lr=0.05, momentum=0.9, nesterov=True, shadow_from=m)
    else:
        o = bvc.FusedAdamW(m.parameters(), lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05, shadow_from=m)
    return m, o, torch.amp.GradScaler("cuda")


def _eager_step(m, o, s, x, mk):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o.zero_grad()
        loss = m(x, bool_masked_pos=mk).loss
    s.scale(loss).backward()
    s.step(o)
    s.update()
    return loss


@pytest.mark.parametrize("opt_name", ["sgd", "adamw"])
def test_graphed_step_equals_eager_loop(opt_name):
    import bvc_b200 as bvc
    cfg = O.make_config("tiny")
    params = O.init_params(cfg, seed=1, perturb=True)
    B, n_steps = 4, 9
    xs = [O.synthetic_clip(B, cfg, seed=10 + i, image_like=True).cuda() for i in range(3)]
    np.random.seed(7)
    mks = [O.batch_tube_masks(B, cfg.grid, 0.5).cuda() for _ in range(3)]
    # eager reference
    m0, o0, s0 = _setup(cfg, params, opt_name)
    ref, ref_grads = [], None
    for i in range(n_steps):
        if i == 6:
            o0.param_groups[0]["lr"] *= 0.5
        ref.append(float(_eager_step(m0, o0, s0, xs[i % 3], mks[i % 3]).detach()))
        if i == 4:
            ref_grads = {k: p.grad.detach().clone() for k, p in m0.named_parameters()}
    # graphed
    m1, o1, s1 = _setup(cfg, params, opt_name)
    step = bvc.GraphedTrainStep(m1, o1, s1, warmup=2)
    got = []
    for i in range(n_steps):
        if i == 6:
            o1.param_groups[0]["lr"] *= 0.5   # an LR schedule step: the graph must be re-captured with the new value
        got.append(float(step(xs[i % 3], mks[i % 3]).detach()))
        if i == 4:
            torch.cuda.synchronize()
            for k, p in m1.named_parameters():   # .grad holds this step's unscaled gradients, as after eager scaler.step
                assert rel_l2(p.grad, ref_grads[k]) <= 1e-4 or float(ref_grads[k].norm()) < 1e-9, k
    torch.cuda.synchronize()
    assert step.captures == 2 and step.replays == n_steps - 2
    for a, b in zip(got, ref):
        assert abs(a - b) <= 2e-5 * abs(b), (got, ref)
    for (k, p), (_, q) in zip(m1.named_parameters(), m0.named_parameters()):
        assert rel_l2(p.detach(), q.detach()) <= 1e-5, k
    m1.check_mask_status()
    # other input shapes: eager fallback (one warning), then back to the graph
    xb = O.synthetic_clip(2, cfg, seed=99, image_like=True).cuda()
    np.random.seed(8)
    mb = O.batch_tube_masks(2, cfg.grid, 0.5).cuda()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        l_small = float(step(xb, mb).detach())
    assert any("runs eagerly" in str(x.message) for x in w) and np.isfinite(l_small)
    l_again = float(step(xs[0], mks[0]).detach())
    assert np.isfinite(l_again) and step.captures == 2 and step.replays == n_steps - 1
    for p in m1.parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all()


def test_graphed_step_rejects_bad_arguments():
    import bvc_b200 as bvc
    cfg = O.make_config("tiny")
    m, o, s = _setup(cfg, O.init_params(cfg, seed=1), "sgd")
    with pytest.raises(ValueError):
        bvc.GraphedTrainStep(m, o, s, warmup=1)
    with pytest.raises(TypeError):
        bvc.GraphedTrainStep(torch.nn.Linear(2, 2), o, s)
