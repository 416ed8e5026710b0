"""FusedSGD (libbvc.so bvc_sgd_step) against torch.optim.SGD -- the reference's optimizer line
(pretrain_videomae.py:187-189) -- on identical parameters and gradients, plain and under GradScaler
(pretrain_videomae.py:312-314), including a skipped (overflow) step.  fp32 arithmetic with the same operation order:
the tolerance is a few ulp per step (fma contraction vs separate mul + add), 2e-6."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(768, 1536), (384,), (3, 5, 7), (1, 1, 384), (2304, 768), (13,)]


def _params(seed, dev):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g).to(dev).requires_grad_(True) for s in SHAPES]


def _grads(seed, dev, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(*s, generator=g) * scale).to(dev) for s in SHAPES]


@pytest.mark.parametrize("kw", [dict(lr=0.1, momentum=0.9, nesterov=True),
                                dict(lr=0.05, momentum=0.8, dampening=0.1, weight_decay=1e-2),
                                dict(lr=0.2, weight_decay=5e-4)])
def test_fused_sgd_matches_torch(kw):
    import bvc_b200 as bvc
    dev = torch.device("cuda:0")
    pa, pb = _params(0, dev), _params(0, dev)
    oa, ob = torch.optim.SGD(pa, **kw), bvc.FusedSGD(pb, **kw)
    for step in range(4):
        for p, q, g in zip(pa, pb, _grads(10 + step, dev)):
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for p, q in zip(pa, pb):
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-6), float((p - q).abs().max())
    if kw.get("momentum", 0):
        for p, q in zip(pa, pb):
            assert torch.allclose(oa.state[p]["momentum_buffer"], ob.state[q]["momentum_buffer"], rtol=2e-6, atol=2e-6)
    assert all(q._version > 0 for q in pb)  # autograd is told about the in-place update
    if kw.get("momentum", 0):  # same per-parameter state keys as torch.optim.SGD (checkpoints interchange)
        assert set(ob.state_dict()["state"][0]) == set(oa.state_dict()["state"][0])


def test_fused_sgd_under_gradscaler_with_overflow_skip():
    import bvc_b200 as bvc
    dev = torch.device("cuda:0")
    kw = dict(lr=0.1, momentum=0.9, nesterov=True)
    pa, pb = _params(1, dev), _params(1, dev)
    oa, ob = torch.optim.SGD(pa, **kw), bvc.FusedSGD(pb, **kw)
    sa, sb = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for step in range(5):
        scale_now = float(sa.get_scale())
        gs = _grads(20 + step, dev, scale=scale_now)
        if step in (0, 2):  # overflow on the very first step (momentum still uninitialised) and later
            gs[1][3] = float("inf")
        for p, q, g in zip(pa, pb, gs):
            p.grad, q.grad = g.clone(), g.clone()
        # the scaler's bookkeeping normally starts in scale(); emulate it
        sa.scale(torch.ones((), device=dev))
        sb.scale(torch.ones((), device=dev))
        sa.step(oa)
        sb.step(ob)
        if step not in (0, 2):
            # .grad holds the UNSCALED gradient after step (loggingtools.py:107-118 reads it).  torch's foreach SGD
            # additionally leaves its in-place nesterov update (grad += momentum * buf) in .grad; the fused kernel
            # writes back the plain unscaled gradient, so compare against that.
            for q, g in zip(pb, gs):
                assert torch.allclose(q.grad, g / scale_now, rtol=2e-6, atol=2e-6)
        sa.update()
        sb.update()
        assert sa.get_scale() == sb.get_scale()
    torch.cuda.synchronize()
    for p, q in zip(pa, pb):
        assert torch.isfinite(q).all()
        assert torch.allclose(p, q, rtol=2e-6, atol=2e-6), float((p - q).abs().max())


def test_fused_sgd_refreshes_weight_copies():
    """shadow_from=model: the bf16 operand copies are rewritten by the optimizer pass, so a second forward after the
    step gives the same loss as a model whose copies were re-cast from scratch."""
    import numpy as np
    import bvc_b200 as bvc
    from oracle import videomae_oracle as O
    from tests.helpers import bvc_config
    dev = torch.device("cuda:0")
    cfg = O.make_config("tiny")
    params = O.init_params(cfg, seed=1, perturb=True)
    x = O.synthetic_clip(2, cfg, seed=2, image_like=True).to(dev)
    np.random.seed(3)
    mask = O.batch_tube_masks(2, cfg.grid, 0.5).to(dev)
    losses = []
    for fused in (False, True):
        model = bvc.VideoMAEForPreTraining(bvc_config(cfg))
        model.load_state_dict(params, strict=True)
        model = model.to(dev).train()
        opt = (bvc.FusedSGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True, shadow_from=model) if fused
               else torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True))
        seq = []
        for _ in range(3):
            opt.zero_grad()
            loss = model(x, bool_masked_pos=mask).loss
            loss.backward()
            opt.step()
            seq.append(float(loss))
        losses.append(seq)
    assert opt.table_builds <= 3  # stable pointers in the real loop: first step, second step (momentum flags), cached
    assert losses[0][1] < losses[0][0]
    for a, b in zip(*losses):
        assert abs(a - b) <= 2e-5 * abs(a), losses


@pytest.mark.parametrize("cls_name,kw", [("AdamW", dict(lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)),   # the reference's
                                         ("AdamW", dict(lr=3e-3, betas=(0.8, 0.99), eps=1e-6, weight_decay=0.0)),
                                         ("Adam", dict(lr=1e-3, betas=(0.9, 0.999), weight_decay=1e-2)),
                                         ("Adam", dict(lr=2e-3))])
def test_fused_adam_matches_torch(cls_name, kw):
    """FusedAdamW / FusedAdam (bvc_adam_step) against torch.optim.AdamW / Adam (pretrain_videomae.py:190-193: AdamW with
    betas (0.9, 0.95)): parameters and both moments after 5 steps.  Same operation order as torch's single-tensor path,
    scalars in double; what differs is fma contraction -- a few ulp per step: rtol 5e-6 / atol 1e-7 (the moments pass
    through zero: absolute floor of about one ulp of their typical magnitude)."""
    import bvc_b200 as bvc
    dev = torch.device("cuda:0")
    pa, pb = _params(2, dev), _params(2, dev)
    oa = getattr(torch.optim, cls_name)(pa, **kw)
    ob = getattr(bvc, "Fused" + cls_name)(pb, **kw)
    for step in range(5):
        for p, q, g in zip(pa, pb, _grads(30 + step, dev)):
            p.grad, q.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    for p, q in zip(pa, pb):
        assert torch.allclose(p, q, rtol=5e-6, atol=1e-7), float((p - q).abs().max())
        for k in ("exp_avg", "exp_avg_sq"):
            assert torch.allclose(oa.state[p][k], ob.state[q][k], rtol=5e-6, atol=1e-7), k
        assert float(ob.state[q]["step"]) == float(oa.state[p]["step"]) == 5.0
    assert set(ob.state_dict()["state"][0]) == set(oa.state_dict()["state"][0])  # checkpoints interchange
    # ... and do: torch's state loads into the fused optimizer and both continue identically
    oc = getattr(bvc, "Fused" + cls_name)(_params(2, dev), **kw)
    with torch.no_grad():
        for p, q in zip(pa, oc.param_groups[0]["params"]):
            q.copy_(p)
    # deep copy: Optimizer.load_state_dict keeps tensors that already have the right dtype / device by reference, and
    # the two optimizers must not share their moment buffers
    oc.load_state_dict(copy.deepcopy(oa.state_dict()))
    for p, q, g in zip(pa, oc.param_groups[0]["params"], _grads(40, dev)):
        p.grad, q.grad = g.clone(), g.clone()
    oa.step()
    oc.step()
    for p, q in zip(pa, oc.param_groups[0]["params"]):
        assert torch.allclose(p, q, rtol=5e-6, atol=1e-7)


@pytest.mark.parametrize("opt_name", ["SGD", "AdamW"])
def test_scaler_step_is_two_launches_and_matches_torch(opt_name):
    """scaler.step(optimizer) through the `grad_scaler` hand-over: the inf / nan check is ONE read-only launch
    (bvc_grad_nonfinite), the update another; same parameters, same scale trajectory and same skipped steps as torch's
    optimizer under the same GradScaler, and a torch SGD state dict saved BEFORE its first step
    (momentum_buffer = None) loads."""
    import bvc_b200 as bvc
    from bvc_b200 import _lib as L
    dev = torch.device("cuda:0")
    kw = dict(lr=0.1, momentum=0.9, nesterov=True) if opt_name == "SGD" else dict(lr=1e-3, betas=(0.9, 0.95), weight_decay=0.05)
    pa, pb = _params(3, dev), _params(3, dev)
    oa = getattr(torch.optim, opt_name)(pa, **kw)
    ob = getattr(bvc, "Fused" + opt_name)(pb, **kw)
    if opt_name == "SGD":
        for p, g in zip(pa, _grads(1, dev)):
            p.grad = g
        sd = torch.optim.SGD(pa, **kw).state_dict()
        # torch's own layout for "momentum not started": the key present with value None
        sd["state"] = {i: {"momentum_buffer": None} for i in range(len(pa))}
        ob.load_state_dict(sd)
    sa, sb = torch.amp.GradScaler("cuda", init_scale=256.0, growth_interval=2), \
        torch.amp.GradScaler("cuda", init_scale=256.0, growth_interval=2)
    for step in range(6):
        scale_now = float(sa.get_scale())
        gs = _grads(50 + step, dev, scale=scale_now)
        if step in (0, 3):
            gs[4][7, 5] = float("nan") if step == 0 else float("-inf")
        for p, q, g in zip(pa, pb, gs):
            p.grad, q.grad = g.clone(), g.clone()
        sa.scale(torch.ones((), device=dev))
        sb.scale(torch.ones((), device=dev))
        sa.step(oa)
        n0 = L.launch_count()
        sb.step(ob)
        assert L.launch_count() - n0 == (2 if opt_name == "SGD" else 3)  # check + update (+ Adam's step counter)
        sa.update()
        sb.update()
        assert sa.get_scale() == sb.get_scale(), step
    torch.cuda.synchronize()
    for p, q in zip(pa, pb):
        assert torch.isfinite(q).all()
        assert torch.allclose(p, q, rtol=5e-6, atol=2e-6), float((p - q).abs().max())
    if opt_name == "AdamW":
        assert float(ob.state[pb[0]]["step"]) == 4.0   # two of the six steps were skipped
