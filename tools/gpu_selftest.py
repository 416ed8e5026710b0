#!/usr/bin/env python
"""Bring-up self-test of every libbvc.so kernel against plain torch references on the GPU box.
Prints one PASS/FAIL line per case and never stops at the first failure (one gpurun call = many answers).
    python tools/gpu_selftest.py [gemm] [rows] [patchify] [attn] [perf]
"""
import math
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bvc_b200  # noqa: E402
from bvc_b200 import _lib as L  # noqa: E402
from oracle import videomae_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
RESULTS = []


def report(name, ok, detail=""):
    RESULTS.append((name, ok))
    print(("PASS " if ok else "FAIL ") + name + ("  " + detail if detail else ""), flush=True)


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def guarded(fn):
    def w(*a, **k):
        try:
            fn(*a, **k)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            report(fn.__name__ + str(a), False, "EXC " + repr(e)[:300])
            traceback.print_exc()
    return w


def bf(t):
    return t.to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ GEMM
@guarded
def gemm_case(M, N, K, a_mn, b_mn, bn, mode="plain", ks=1, pair=0):
    g = torch.Generator(device=dev).manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A = bf(torch.randn(M, K, device=dev, generator=g))
    Bm = bf(torch.randn(N, K, device=dev, generator=g) * 0.5)
    a_store = A.t().contiguous() if a_mn else A
    b_store = Bm.t().contiguous() if b_mn else Bm
    ref = A.double() @ Bm.double().t()
    name = f"gemm M{M} N{N} K{K} a_mn{a_mn} b_mn{b_mn} bn{bn} {mode} ks{ks}" + (" pair" if pair == 2 else "")
    kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=bn, cta_pair=pair)
    if mode == "plain":
        of = torch.full((M, N), float("nan"), device=dev)
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.gemm(a_store, b_store, M, N, K, out_f32=of, out_bf16=ob, **kw)
        e1, e2 = relerr(of, ref), relerr(ob.float(), ref)
        report(name, e1 < 1e-5 and e2 < 5e-3, f"f32 {e1:.2e} bf16 {e2:.2e}")
    elif mode == "splitk":
        of = torch.zeros(M, N, device=dev)
        L.gemm(a_store, b_store, M, N, K, out_f32=of, k_splits=ks, alpha=0.5, **kw)
        e1 = relerr(of, 0.5 * ref)
        report(name, e1 < 1e-5, f"f32 {e1:.2e}")
    elif mode == "bias_gelu":
        bias = torch.randn(N, device=dev, generator=g)
        aux = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.gemm(a_store, b_store, M, N, K, out_bf16=ob, bias=bias, act=1, aux_out=aux, ld_aux=N, alpha=0.05, **kw)
        pre = bf((0.05 * ref + bias.double()).float())
        out = torch.nn.functional.gelu(pre.float())
        xg = pre.double().requires_grad_(True)
        torch.nn.functional.gelu(xg).sum().backward()   # aux = gelu'(pre), what the backward GEMM multiplies by
        e1, e2 = relerr(aux.float(), xg.grad), relerr(ob.float(), out)
        report(name, e1 < 3e-3 and e2 < 5e-3, f"aux {e1:.2e} out {e2:.2e}")
    elif mode == "gelu_bwd":
        gp = bf(torch.randn(M, N, device=dev, generator=g))   # the saved gelu' factor
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.gemm(a_store, b_store, M, N, K, out_bf16=ob, act=2, aux_in=gp, ld_aux=N, alpha=0.05, **kw)
        e = relerr(ob.float(), 0.05 * ref * gp.double())
        report(name, e < 5e-3, f"{e:.2e}")
    elif mode == "residual_idx_seg":
        # out rows remapped into segments, residual rows gathered through an index (enc->dec + pos[vis_idx])
        seg, stride, off = 5, 12, 3
        nb = (M + seg - 1) // seg
        tab = torch.randn(97, N, device=dev, generator=g)
        idx = torch.randint(0, 97, (M,), device=dev, generator=g, dtype=torch.int32)
        of = torch.zeros(nb * stride, N, device=dev)
        L.gemm(a_store, b_store, M, N, K, out_f32=of, res=tab, ldr=N, res_idx=idx, out_seg=seg, out_seg_stride=stride,
               out_seg_off=off, **kw)
        want = torch.zeros_like(of, dtype=torch.float64)
        r = torch.arange(M, device=dev)
        want[(r // seg) * stride + r % seg + off] = ref + tab[idx.long()].double()
        e = relerr(of, want)
        report(name, e < 1e-5, f"{e:.2e}")
    elif mode == "residual":
        res = torch.randn(M, N, device=dev, generator=g)
        bias = torch.randn(N, device=dev, generator=g)
        of = torch.zeros(M, N, device=dev)
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.gemm(a_store, b_store, M, N, K, out_f32=of, out_bf16=ob, res=res, ldr=N, bias=bias, **kw)
        want = ref + res.double() + bias.double()
        e1, e2 = relerr(of, want), relerr(ob.float(), want)
        report(name, e1 < 1e-5 and e2 < 5e-3, f"f32 {e1:.2e} bf16 {e2:.2e}")
    elif mode == "bias_bf16":
        # the engine's QKV projection (EPI_PLAIN: bf16 output only, bias)
        bias = torch.randn(N, device=dev, generator=g)
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.gemm(a_store, b_store, M, N, K, out_bf16=ob, bias=bias, **kw)
        e2 = relerr(ob.float(), ref + bias.double())
        report(name, e2 < 5e-3, f"bf16 {e2:.2e}")
    elif mode == "res_f32":
        # exactly the engine's residual GEMMs (EPI_RES: fp32 output only, bias + fp32 residual rows)
        res = torch.randn(M, N, device=dev, generator=g)
        bias = torch.randn(N, device=dev, generator=g)
        of = torch.zeros(M, N, device=dev)
        L.gemm(a_store, b_store, M, N, K, out_f32=of, res=res, ldr=N, bias=bias, **kw)
        e1 = relerr(of, ref + res.double() + bias.double())
        report(name, e1 < 1e-5, f"f32 {e1:.2e}")
    elif mode == "gelu_bwd_colsum":
        # the engine's fc2 dgrad: x gelu' (saved factor) with the fc1 bias gradient (column sums) fused
        gp = bf(torch.randn(M, N, device=dev, generator=g))
        ob = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        cs = torch.zeros(N, device=dev)
        L.gemm(a_store, b_store, M, N, K, out_bf16=ob, act=2, aux_in=gp, ld_aux=N, alpha=0.05, colsum=cs, **kw)
        want = 0.05 * ref * gp.double()
        e, ec = relerr(ob.float(), want), relerr(cs, want.sum(0))
        report(name, e < 5e-3 and ec < 2e-3, f"{e:.2e} colsum {ec:.2e}")
    elif mode == "loss_nologits":
        tgt = torch.randn(M, N, device=dev, generator=g)
        bias = torch.randn(N, device=dev, generator=g)
        part = torch.full((L.gemm_loss_slots(M, N, bn),), float("nan"), device=dev)
        diff = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.gemm(a_store, b_store, M, N, K, out_bf16=diff, bias=bias, target=tgt, ldt=N, loss_partial=part, alpha=0.1, **kw)
        loss = torch.zeros(1, device=dev)
        L.loss_finalize(part, M * N, torch.zeros(1, device=dev, dtype=torch.int32), loss)
        lg = 0.1 * ref + bias.double()
        want = ((lg - tgt.double()) ** 2).mean()
        e0, e1 = abs(float(loss) - float(want)) / float(want), relerr(diff.float(), lg - tgt.double())
        report(name, e0 < 1e-5 and e1 < 5e-3, f"loss {e0:.2e} diff {e1:.2e}")
    elif mode == "loss":
        tgt = torch.randn(M, N, device=dev, generator=g)
        bias = torch.randn(N, device=dev, generator=g)
        slots = L.gemm_loss_slots(M, N, bn)
        part = torch.full((slots,), float("nan"), device=dev)
        diff = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        logits = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        L.gemm(a_store, b_store, M, N, K, out_bf16=diff, bias=bias, target=tgt, ldt=N, loss_partial=part,
               logits_out=logits, alpha=0.1, **kw)
        loss = torch.zeros(1, device=dev)
        status = torch.zeros(1, device=dev, dtype=torch.int32)
        L.loss_finalize(part, M * N, status, loss)
        lg = 0.1 * ref + bias.double()
        want = ((lg - tgt.double()) ** 2).mean()
        e0 = abs(float(loss) - float(want)) / float(want)
        e1, e2 = relerr(diff.float(), lg - tgt.double()), relerr(logits.float(), lg)
        report(name, e0 < 1e-5 and e1 < 5e-3 and e2 < 5e-3, f"loss {e0:.2e} diff {e1:.2e} logits {e2:.2e}")


def run_gemm_major(a_mn, b_mn):
    gemm_case(128, 128, 64, a_mn, b_mn, 128)
    gemm_case(128, 64, 256, a_mn, b_mn, 64)
    gemm_case(256, 256, 128, a_mn, b_mn, 256)
    gemm_case(384, 192, 192, a_mn, b_mn, 192)
    gemm_case(200, 136, 72, a_mn, b_mn, 0)  # ragged M / N / K tails
    gemm_case(1000, 768, 1536, a_mn, b_mn, 0)


def run_gemm_epilogues():
    gemm_case(640, 2304, 768, 0, 0, 256)
    gemm_case(4096, 384, 1536, 0, 1, 0)
    gemm_case(768, 1536, 4096, 1, 1, 0, "splitk", 0)
    gemm_case(384, 1152, 2048, 1, 1, 128, "splitk", 5)
    gemm_case(256, 256, 640, 0, 0, 128, "splitk", 3)
    for bn in (64, 128, 192, 256):
        gemm_case(300, 3 * bn, 256, 0, 0, bn, "bias_gelu")
    gemm_case(300, 512, 256, 0, 1, 0, "gelu_bwd")
    gemm_case(300, 512, 256, 0, 0, 0, "residual")
    gemm_case(300, 384, 128, 0, 0, 0, "residual_idx_seg")
    gemm_case(300, 1536, 384, 0, 0, 0, "loss")
    gemm_case(2816, 1536, 384, 0, 0, 256, "loss")


def run_gemm_pair():
    """CTA-pair (cta_group::2, 256 x BN) tiles: every operand major, every fused epilogue, ragged edges (a pair whose
    second CTA lies entirely past M), multi-tile persistence and split-K."""
    gemm_case(256, 256, 64, 0, 0, 256, pair=2)
    gemm_case(256, 128, 256, 0, 0, 128, pair=2)
    gemm_case(512, 384, 192, 0, 0, 192, pair=2)
    for a_mn, b_mn in ((0, 0), (0, 1), (1, 0), (1, 1)):
        gemm_case(200, 136, 72, a_mn, b_mn, 128, pair=2)     # ragged M / N / K tails
        gemm_case(384, 512, 200, a_mn, b_mn, 256, pair=2)    # second CTA of the last pair is past M
        gemm_case(1000, 768, 1536, a_mn, b_mn, 256, pair=2)
        gemm_case(40000, 768, 512, a_mn, b_mn, 256, pair=2)  # several tiles per cluster (both accumulator stages)
    gemm_case(1000, 768, 1536, 0, 0, 192, pair=2)
    gemm_case(768, 1536, 4096, 1, 1, 256, "splitk", 0, pair=2)
    gemm_case(384, 1152, 2048, 1, 1, 128, "splitk", 5, pair=2)
    for bn in (128, 192, 256):
        gemm_case(300, 3 * bn, 256, 0, 0, bn, "bias_gelu", pair=2)
        gemm_case(2560, 3 * bn, 256, 0, 0, bn, "bias_gelu", pair=2)
    gemm_case(2560, 512, 256, 0, 1, 256, "gelu_bwd", pair=2)
    gemm_case(2560, 512, 256, 0, 0, 256, "residual", pair=2)
    gemm_case(300, 384, 128, 0, 0, 128, "residual_idx_seg", pair=2)
    gemm_case(300, 1536, 384, 0, 0, 256, "loss", pair=2)
    gemm_case(2816, 1536, 384, 0, 0, 256, "loss", pair=2)


# ------------------------------------------------------------------------------------------------ rows
@guarded
def ln_case(M, d, seg=(0, 0, 0)):
    g = torch.Generator(device=dev).manual_seed(M + d)
    rows_phys = M if seg[0] == 0 else ((M + seg[0] - 1) // seg[0]) * seg[1]
    x = torch.randn(rows_phys, d, device=dev, generator=g) * 2 + 0.5
    gamma = torch.randn(d, device=dev, generator=g)
    beta = torch.randn(d, device=dev, generator=g)
    r = torch.arange(M, device=dev)
    pr = r if seg[0] == 0 else (r // seg[0]) * seg[1] + r % seg[0] + seg[2]
    y = torch.zeros(M, d, device=dev, dtype=torch.bfloat16)
    mean = torch.zeros(M, device=dev)
    rstd = torch.zeros(M, device=dev)
    L.layernorm_fwd(x, gamma, beta, 1e-12, M, d, y, mean, rstd, seg=seg)
    xr = x[pr].double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (d,), gd, bd, 1e-12)
    e = relerr(y.float(), ref)
    em = relerr(mean, xr.mean(1))
    report(f"ln_fwd M{M} d{d} seg{seg}", e < 4e-3 and em < 1e-5, f"y {e:.2e} mean {em:.2e}")
    dy = bf(torch.randn(M, d, device=dev, generator=g))
    dres = torch.randn(rows_phys, d, device=dev, generator=g)
    dxf = torch.zeros(rows_phys, d, device=dev)
    dxb = torch.zeros(rows_phys, d, device=dev, dtype=torch.bfloat16)
    dg = torch.zeros(d, device=dev)
    db = torch.zeros(d, device=dev)
    dsum = torch.zeros(d, device=dev)
    L.layernorm_bwd(dy, x, mean, rstd, gamma, dres, M, d, dxf, dxb, dg, db, seg=seg, dxsum=dsum)
    ref.backward(dy.double())
    want = xr.grad + dres[pr].double()
    e1, e2 = relerr(dxf[pr], want), relerr(dxb[pr].float(), want)
    e3, e4, e5 = relerr(dg, gd.grad), relerr(db, bd.grad), relerr(dsum, want.sum(0))
    report(f"ln_bwd M{M} d{d} seg{seg}", e1 < 1e-5 and e2 < 4e-3 and e3 < 1e-4 and e4 < 1e-4 and e5 < 1e-4,
           f"dx {e1:.2e} dxb {e2:.2e} dgamma {e3:.2e} dbeta {e4:.2e} dxsum {e5:.2e}")


@guarded
def colsum_case(M, N, f32, seg=(0, 0, 0)):
    g = torch.Generator(device=dev).manual_seed(M + N)
    rows_phys = M if seg[0] == 0 else ((M + seg[0] - 1) // seg[0]) * seg[1]
    x = torch.randn(rows_phys, N, device=dev, generator=g)
    if not f32:
        x = bf(x)
    r = torch.arange(M, device=dev)
    pr = r if seg[0] == 0 else (r // seg[0]) * seg[1] + r % seg[0] + seg[2]
    out = torch.zeros(N, device=dev)
    sd = torch.full((1,), 3.0, device=dev)
    L.colsum(x, M, N, out, seg=seg, scale=0.5, scale_dev=sd)
    want = 1.5 * x[pr].double().sum(0)
    e = relerr(out, want)
    report(f"colsum M{M} N{N} f32={f32} seg{seg}", e < 1e-5, f"{e:.2e}")


@guarded
def misc_rows():
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn(1000003, device=dev, generator=g)
    y = torch.zeros(1000003, device=dev, dtype=torch.bfloat16)
    L.cast_bf16(x, y)
    report("cast_bf16", torch.equal(y, bf(x)))
    B, N, nv, d = 3, 40, 7, 64
    src = torch.randn(B * N, d, device=dev, generator=g)
    dst = torch.zeros(B * nv, d, device=dev, dtype=torch.bfloat16)
    L.rows_to_bf16(src, B * nv, d, dst, seg=(nv, N, 0))
    report("rows_to_bf16 seg", torch.equal(dst.view(B, nv, d), bf(src.view(B, N, d)[:, :nv])))
    pos = torch.randn(N, d, device=dev, generator=g)
    tok = torch.randn(d, device=dev, generator=g)
    msk = torch.stack([torch.randperm(N, device=dev, generator=g)[: N - nv].sort().values for _ in range(B)]).int()
    xx = torch.zeros(B, N, d, device=dev)
    L.decoder_mask_rows(xx, tok, pos, msk, B, N, nv, d)
    want = torch.zeros_like(xx)
    want[:, nv:] = tok + pos[msk.long()]
    report("decoder_mask_rows", torch.equal(xx, want))


def run_rows():
    for M, d in ((1000, 768), (777, 384), (64, 64), (300, 1024), (513, 192), (100, 512)):
        ln_case(M, d)
    ln_case(5 * 11, 384, seg=(11, 20, 9))
    colsum_case(10240, 768, False)
    colsum_case(1000, 1536, False)
    colsum_case(777, 384, True)
    colsum_case(5 * 11, 384, True, seg=(11, 20, 9))
    misc_rows()


# ------------------------------------------------------------------------------------------------ mask + patchify
@guarded
def patchify_case(B, cfgname, ratio, norm_pix=True, **over):
    cfg = O.make_config(cfgname, **over)
    x = O.synthetic_clip(B, cfg, seed=B, image_like=True)
    np.random.seed(B)
    mask = O.batch_tube_masks(B, cfg.grid, ratio)
    vis_ref, msk_ref = O.mask_to_index(mask)
    N = cfg.seq_len
    nv = vis_ref.shape[1]
    mg = mask.to(dev).to(torch.uint8)
    cnt = torch.zeros(B, device=dev, dtype=torch.int32)
    L.mask_count(mg, cnt)
    vis = torch.zeros(B, nv, device=dev, dtype=torch.int32)
    msk = torch.zeros(B, N - nv, device=dev, dtype=torch.int32)
    slot = torch.zeros(B, N, device=dev, dtype=torch.int32)
    status = torch.zeros(1, device=dev, dtype=torch.int32)
    L.mask_to_index(mg, nv, vis, msk, slot, status)
    ok = (bool((cnt.cpu() == nv).all()) and torch.equal(vis.cpu(), vis_ref) and torch.equal(msk.cpu(), msk_ref)
          and int(status) == 0)
    report(f"mask_to_index B{B} {cfgname} N{N} nv{nv}", ok)
    K = cfg.patch_dim
    pv = torch.zeros(B * nv, K, device=dev, dtype=torch.bfloat16)
    tgt = torch.full((B * (N - nv), K), float("nan"), device=dev)
    xg = x.to(dev)
    L.patchify_target(xg, slot, cfg.tubelet_size, cfg.patch_size, nv, pv, tgt, norm_pix)
    torch.cuda.synchronize()
    cfg.norm_pix_loss = norm_pix
    pe = O.patchify_embed_order(x, cfg)
    pv_ref = torch.gather(pe, 1, vis_ref.long()[:, :, None].expand(-1, -1, K))
    ok1 = torch.equal(pv.cpu().view(B, nv, K), pv_ref.to(torch.bfloat16))
    t_ref = O.norm_pix_target(x, msk_ref, cfg)
    err = float((tgt.cpu().view(B, N - nv, K) - t_ref).abs().max())
    report(f"patchify_target B{B} {cfgname} norm={norm_pix}", ok1 and err < 2e-5, f"patch bit-exact {ok1}, target maxabs {err:.2e}")


@guarded
def bad_mask_case():
    B, N = 4, 64
    m = torch.zeros(B, N, dtype=torch.uint8)
    m[:, :32] = 1
    m[2, 40] = 1
    mg = m.to(dev)
    vis = torch.zeros(B, 32, device=dev, dtype=torch.int32)
    msk = torch.zeros(B, 32, device=dev, dtype=torch.int32)
    slot = torch.zeros(B, N, device=dev, dtype=torch.int32)
    status = torch.zeros(1, device=dev, dtype=torch.int32)
    L.mask_to_index(mg, 32, vis, msk, slot, status)
    report("mask_to_index flags unequal rows", int(status) == 1)


def run_patchify():
    patchify_case(3, "tiny", 0.5)
    patchify_case(2, "tiny", 0.5, norm_pix=False)
    patchify_case(2, "small", 0.9)
    patchify_case(2, "small", 0.9, num_frames=1, tubelet_size=1)
    bad_mask_case()


# ------------------------------------------------------------------------------------------------ attention
@guarded
def attn_case(B, S, H, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed + S)
    d = H * 64
    qkv = bf(torch.randn(B, S, 3, H, 64, device=dev, generator=g))
    out = torch.zeros(B, S, d, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, S, device=dev)
    scale = 64 ** -0.5
    L.attn_fwd(qkv, B, S, H, scale, out, lse)
    do = bf(torch.randn(B, S, d, device=dev, generator=g))
    dqkv = torch.zeros_like(qkv)
    delta = torch.zeros(B, H, S, device=dev)
    L.attn_bwd(qkv, out, do, lse, B, S, H, scale, delta, dqkv)
    # S > 160: the same call with the fp32 dQ workspace runs the ONE-pass kernel (attn_bwd1); poisoned workspace
    dqkv1 = None
    if S > 160:
        dqkv1 = torch.full_like(qkv, float("nan"))
        L.attn_bwd(qkv, out, do, lse, B, S, H, scale, torch.zeros(B, H, S, device=dev), dqkv1,
                   torch.full((B, S, H, 64), float("nan"), device=dev))
    # fp64 reference, a few clips at a time (the scores of B64 S1568 H6 are 7.5 GB in fp64)
    chunk = max(1, min(B, int(2e9 // (H * S * S * 8))))
    num = {k: 0.0 for k in ("out", "lse", "dq", "dk", "dv") + (("dq1", "dk1", "dv1") if dqkv1 is not None else ())}
    den = dict(num)
    for b0 in range(0, B, chunk):
        sl = slice(b0, min(B, b0 + chunk))
        q, k, v = (qkv[sl, :, i].permute(0, 2, 1, 3).double().requires_grad_(True) for i in range(3))
        s = (q @ k.transpose(-1, -2)) * scale
        ref = torch.softmax(s, -1) @ v
        ref_lse = torch.logsumexp(s, -1)
        del s
        ref.backward(do[sl].double().view(-1, S, H, 64).permute(0, 2, 1, 3))
        pairs = {"out": (out[sl].float().view(-1, S, H, 64).permute(0, 2, 1, 3), ref.detach()), "lse": (lse[sl], ref_lse),
                 "dq": (dqkv[sl, :, 0].permute(0, 2, 1, 3).float(), q.grad),
                 "dk": (dqkv[sl, :, 1].permute(0, 2, 1, 3).float(), k.grad),
                 "dv": (dqkv[sl, :, 2].permute(0, 2, 1, 3).float(), v.grad)}
        if dqkv1 is not None:
            for i_, n_ in enumerate(("dq1", "dk1", "dv1")):
                pairs[n_] = (dqkv1[sl, :, i_].permute(0, 2, 1, 3).float(), (q, k, v)[i_].grad)
        for kk, (a, r) in pairs.items():
            num[kk] += float((a.double() - r.double()).pow(2).sum())
            den[kk] += float(r.double().pow(2).sum())
        del q, k, v, ref, ref_lse, pairs
    err = {kk: (num[kk] / max(den[kk], 1e-60)) ** 0.5 for kk in num}
    report(f"attn_fwd B{B} S{S} H{H}", err["out"] < 6e-3 and err["lse"] < 1e-4, f"out {err['out']:.2e} lse {err['lse']:.2e}")
    report(f"attn_bwd B{B} S{S} H{H}", max(err["dq"], err["dk"], err["dv"]) < 1.5e-2,
           f"dq {err['dq']:.2e} dk {err['dk']:.2e} dv {err['dv']:.2e}")
    if dqkv1 is not None:
        report(f"attn_bwd one-pass B{B} S{S} H{H}", max(err["dq1"], err["dk1"], err["dv1"]) < 1.5e-2,
               f"dq {err['dq1']:.2e} dk {err['dk1']:.2e} dv {err['dv1']:.2e}")


def run_attn():
    attn_case(1, 128, 1)
    attn_case(2, 160, 2)
    attn_case(2, 8, 1)
    attn_case(1, 1568, 2)
    attn_case(3, 200, 3)
    attn_case(2, 392, 6)
    # short sequences: the whole-sequence-resident kernels of attn_small.cu (S <= 192), many items per CTA
    attn_case(3, 192, 2)
    attn_case(2, 144, 3)
    attn_case(2, 100, 1)
    attn_case(2, 129, 2)
    attn_case(40, 160, 12)
    attn_case(2, 2, 1)   # (S = 1 has an exactly-zero dq / dk reference: no relative error to take)
    attn_case(2, 17, 2)
    attn_case(1, 33, 1)
    attn_case(3, 159, 2)


def run_bench_shapes():
    """The kernel shapes of the BENCHMARK step (ViT-B/16, batch 64: M = 64 x 160 = 10240 encoder rows, 64 x 1568 =
    100352 decoder rows, 64 x 1408 = 90112 masked rows) with the tile shapes the dispatcher picks for them on its own
    (block_n = 0, cta_pair = 0): persistent schedulers with > 148 tiles, CTA-pair auto-selection, both accumulator
    stages, the fused epilogues exactly as engine.py calls them."""
    gemm_case(100352, 1536, 384, 0, 0, 0, "bias_gelu")          # decoder fc1
    gemm_case(100352, 1536, 384, 0, 1, 0, "gelu_bwd_colsum")    # decoder fc2 dgrad x gelu'
    gemm_case(100352, 384, 1536, 0, 0, 0, "res_f32")            # decoder fc2 (CTA pairs)
    gemm_case(100352, 384, 384, 0, 0, 0, "res_f32")             # decoder out-proj
    gemm_case(100352, 1152, 384, 0, 0, 0, "bias_bf16")          # decoder qkv
    gemm_case(10240, 2304, 768, 0, 0, 0, "bias_bf16")           # encoder qkv
    gemm_case(100352, 384, 1152, 0, 1, 0, "bias_bf16")          # decoder qkv dgrad (B MN-major)
    gemm_case(90112, 1536, 384, 0, 0, 192, "loss_nologits")     # head + MSE
    gemm_case(1536, 384, 100352, 1, 1, 0, "splitk", 0)          # decoder fc1 wgrad
    gemm_case(10240, 3072, 768, 0, 0, 0, "bias_gelu")           # encoder fc1
    gemm_case(10240, 3072, 768, 0, 1, 0, "gelu_bwd_colsum")     # encoder fc2 dgrad
    gemm_case(10240, 768, 3072, 0, 0, 0, "res_f32")             # encoder fc2 (CTA pairs)
    gemm_case(10240, 768, 768, 0, 0, 0, "res_f32")              # encoder out-proj
    gemm_case(3072, 768, 10240, 1, 1, 0, "splitk", 0)           # encoder fc1 wgrad
    attn_case(64, 160, 12)                                      # encoder attention (short-sequence kernels)
    attn_case(64, 1568, 6)                                      # decoder attention
    patchify_case(64, "base", 0.9)
    ln_case(10240, 768)
    ln_case(64 * 1408, 384, seg=(1408, 1568, 160))              # final decoder norm over the masked rows


# ------------------------------------------------------------------------------------------------ perf probes
def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


@guarded
def run_perf():
    shapes = [  # (M, N, K, a_mn, b_mn, note)
        (10240, 2304, 768, 0, 0, "enc qkv"), (10240, 768, 768, 0, 0, "enc proj"), (10240, 3072, 768, 0, 0, "enc fc1"),
        (10240, 768, 3072, 0, 0, "enc fc2"), (100352, 1152, 384, 0, 0, "dec qkv"), (100352, 384, 384, 0, 0, "dec proj"),
        (100352, 1536, 384, 0, 0, "dec fc1"), (100352, 384, 1536, 0, 0, "dec fc2"), (90112, 1536, 384, 0, 0, "head"),
        (100352, 384, 1536, 0, 1, "dec fc1 dgrad"), (1536, 384, 100352, 1, 1, "dec fc1 wgrad"),
        (3072, 768, 10240, 1, 1, "enc fc1 wgrad"), (8192, 8192, 8192, 0, 0, "square"),
    ]
    for M, N, K, a_mn, b_mn, note in shapes:
        A = bf(torch.randn(K if a_mn else M, M if a_mn else K, device=dev))
        Bm = bf(torch.randn(K if b_mn else N, N if b_mn else K, device=dev))
        wg = a_mn and b_mn
        of = torch.zeros(M, N, device=dev) if wg else None
        ob = None if wg else torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        for bn in (0, 128, 256):
            ms = timeit(lambda: L.gemm(A, Bm, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), out_f32=of, out_bf16=ob,
                                       k_splits=0 if wg else 1, block_n=bn))
            print(f"PERF gemm {note:14s} M{M} N{N} K{K} bn{bn}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
        At, Bt = (A.t() if a_mn else A), (Bm.t() if b_mn else Bm)
        ms = timeit(lambda: torch.matmul(At, Bt.t()))
        print(f"PERF torch {note:14s}: {ms*1e3:8.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)
    # patchify at the headline shape
    cfg = O.make_config("base")
    Bc = 64
    x = torch.randn(Bc, 16, 3, 224, 224, device=dev)
    np.random.seed(0)
    mask = O.batch_tube_masks(Bc, cfg.grid, 0.9).to(dev).to(torch.uint8)
    nv, N = 160, 1568
    vis = torch.zeros(Bc, nv, device=dev, dtype=torch.int32)
    msk = torch.zeros(Bc, N - nv, device=dev, dtype=torch.int32)
    slot = torch.zeros(Bc, N, device=dev, dtype=torch.int32)
    status = torch.zeros(1, device=dev, dtype=torch.int32)
    L.mask_to_index(mask, nv, vis, msk, slot, status)
    pv = torch.zeros(Bc * nv, 1536, device=dev, dtype=torch.bfloat16)
    tgt = torch.zeros(Bc * (N - nv), 1536, device=dev)
    ms = timeit(lambda: L.patchify_target(x, slot, 2, 16, nv, pv, tgt, True))
    by = x.numel() * 4 + pv.numel() * 2 + tgt.numel() * 4
    print(f"PERF patchify_target B64: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s", flush=True)
    xx = torch.randn(100352, 384, device=dev)
    gam, bet = torch.ones(384, device=dev), torch.zeros(384, device=dev)
    y = torch.zeros(100352, 384, device=dev, dtype=torch.bfloat16)
    mean, rstd = torch.zeros(100352, device=dev), torch.zeros(100352, device=dev)
    ms = timeit(lambda: L.layernorm_fwd(xx, gam, bet, 1e-12, 100352, 384, y, mean, rstd))
    print(f"PERF ln_fwd 100352x384: {ms*1e3:.1f} us  {(xx.numel()*6)/ms/1e6:.0f} GB/s", flush=True)
    dxf = torch.zeros_like(xx)
    dxb = torch.zeros_like(y)
    dg, db = torch.zeros(384, device=dev), torch.zeros(384, device=dev)
    ms = timeit(lambda: L.layernorm_bwd(y, xx, mean, rstd, gam, xx, 100352, 384, dxf, dxb, dg, db))
    print(f"PERF ln_bwd 100352x384: {ms*1e3:.1f} us  {(xx.numel()*16)/ms/1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["rows", "patchify", "gemm00", "gemm01", "gemm10", "gemm11", "gemmx"]
    print("device:", torch.cuda.get_device_name(0), "lib ABI", L.load().bvc_abi_version(), flush=True)
    t0 = time.time()
    for w in which:
        {"gemm00": lambda: run_gemm_major(0, 0), "gemm01": lambda: run_gemm_major(0, 1),
         "gemm10": lambda: run_gemm_major(1, 0), "gemm11": lambda: run_gemm_major(1, 1),
         "gemmx": run_gemm_epilogues, "gemmpair": run_gemm_pair, "benchshapes": run_bench_shapes, "rows": run_rows, "patchify": run_patchify, "attn": run_attn, "perf": run_perf}[w]()
    torch.cuda.synchronize()
    nfail = sum(1 for _, ok in RESULTS if not ok)
    print(f"SELFTEST {len(RESULTS) - nfail}/{len(RESULTS)} passed in {time.time() - t0:.1f}s", flush=True)
    sys.exit(1 if nfail else 0)
