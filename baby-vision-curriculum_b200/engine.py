"""Autograd plumbing of the VideoMAE pretraining step on libbvc.so kernels.

One torch.autograd.Function per stage (embed, each transformer block, encoder_to_decoder, head+loss) so that
parameter gradients become available layer by layer in reverse order -- PyTorch DDP's bucketed all-reduce
(pretrain_videomae.py:180) then overlaps with the rest of the backward exactly as it does for the reference model.
torch is used for memory (torch.empty / zeros), streams and autograd bookkeeping only; every FLOP and every byte of
activation traffic is a libbvc.so kernel (see include/bvc.h).  There is no fallback path.

Dtype flow (SURVEY.md section 9.4 for the reference's): residual stream fp32 in encoder and decoder, GEMM / attention
operands bf16, accumulation fp32, LayerNorm / softmax statistics fp32, master weights and gradients fp32.
"""
from __future__ import annotations

import torch

from . import _lib as L

BF16 = torch.bfloat16
F32 = torch.float32


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


class Bf16Cache:
    """bf16 copies of the fp32 master weights (what autocast re-casts on every call, HF linear layers), refreshed
    only when a parameter's storage or version counter changes (i.e. after an optimizer step / load_state_dict)."""

    def __init__(self):
        self._store = {}

    def _key(self, params):
        return tuple((p.data_ptr(), p._version) for p in params)

    def weight(self, name, p):
        ent = self._store.get(name)
        key = self._key((p,))
        if ent is None or ent[0] != key or ent[1].device != p.device:
            buf = ent[1] if ent is not None and ent[1].device == p.device and ent[1].numel() == p.numel() else \
                _empty((p.numel(),), BF16, p.device)
            L.cast_bf16(p.detach().reshape(-1), buf)
            ent = (key, buf)
            self._store[name] = ent
        return ent[1]

    def qkv(self, name, wq, wk, wv, qb, vb):
        """packed [3d, d] bf16 weight and [3d] fp32 bias (q_bias, 0, v_bias) -- HF:239-242 as one GEMM."""
        ent = self._store.get(name)
        key = self._key((wq, wk, wv, qb, vb))
        if ent is None or ent[0] != key or ent[1].device != wq.device:
            d = wq.shape[0]
            if ent is not None and ent[1].device == wq.device:
                w, b = ent[1], ent[2]
            else:
                w = _empty((3 * d * d,), BF16, wq.device)
                b = torch.zeros(3 * d, dtype=F32, device=wq.device)
            for i, p in enumerate((wq, wk, wv)):
                L.cast_bf16(p.detach().reshape(-1), w[i * d * d:(i + 1) * d * d])
            b[0:d].copy_(qb.detach())
            b[2 * d:3 * d].copy_(vb.detach())
            ent = (key, w, b)
            self._store[name] = ent
        return ent[1], ent[2]

    def clear(self):
        self._store.clear()


class GradSideChannel:
    """The bf16 twin of the most recent fp32 activation gradient (written by the same LayerNorm-backward kernel),
    handed to the upstream stage so it does not have to re-cast its incoming gradient."""

    def __init__(self):
        self.ptr = None
        self.t = None

    def put(self, g_f32, g_bf16):
        self.ptr, self.t = g_f32.data_ptr(), g_bf16

    def drop(self):
        self.ptr, self.t = None, None

    def take(self, g_f32, M, d):
        if self.ptr == g_f32.data_ptr() and self.t is not None and self.t.numel() == M * d:
            t, self.t, self.ptr = self.t, None, None
            return t
        self.t, self.ptr = None, None
        out = _empty((M, d), BF16, g_f32.device)
        L.rows_to_bf16(g_f32, M, d, out)
        return out


class StepState:
    """Per-forward shared state (index lists, shapes, caches)."""

    def __init__(self, cache: Bf16Cache):
        self.cache = cache
        self.side = GradSideChannel()
        self.B = self.N = self.nv = self.nm = 0
        self.vis_idx = self.msk_idx = self.slot = self.status = None


def _contig_grad(g):
    return g if g.is_contiguous() else g.contiguous()


# ==================================================================================================== embed
class EmbedFn(torch.autograd.Function):
    """e = W_pe . patch + b_pe + pos[vis_idx]  on the visible tubelets only (HF:164-177, HF:109-122)."""

    @staticmethod
    def forward(ctx, w, b, patches, pos, st: StepState, name):
        D = w.shape[0]
        K = w.numel() // D
        M = patches.shape[0]
        wb = st.cache.weight(name, w)
        x = _empty((M, D), F32, patches.device)
        L.gemm(patches, wb, M, D, K, out_f32=x, bias=b.detach(), res=pos, ldr=D, res_idx=st.vis_idx)
        ctx.st, ctx.patches, ctx.dims, ctx.wshape = st, patches, (M, D, K), w.shape
        return x

    @staticmethod
    def backward(ctx, dx):
        st, patches = ctx.st, ctx.patches
        M, D, K = ctx.dims
        dx = _contig_grad(dx)
        dxb = st.side.take(dx, M, D)
        flat = torch.zeros(D * K + D, dtype=F32, device=dx.device)
        dw, db = flat[:D * K].view(ctx.wshape), flat[D * K:]
        L.gemm(dxb, patches, D, K, M, a_mn=True, b_mn=True, lda=D, ldb=K, out_f32=dw, k_splits=0)
        L.colsum(dxb, M, D, db)
        return dw, db, None, None, None, None


# ==================================================================================================== block
class BlockFn(torch.autograd.Function):
    """One pre-LN transformer block (HF:348-366): x + Wo.Attn(LN1(x)) then + W2.gelu(W1.LN2(.))."""

    @staticmethod
    def forward(ctx, x, ln1w, ln1b, wq, wk, wv, qb, vb, wo, bo, ln2w, ln2b, w1, b1, w2, b2, st: StepState, name, B, S,
                H, eps):
        dev = x.device
        M, d = x.shape
        ff = w1.shape[0]
        cache = st.cache
        wqkv, bqkv = cache.qkv(name + "qkv", wq, wk, wv, qb, vb)
        wob, w1b, w2b = cache.weight(name + "wo", wo), cache.weight(name + "w1", w1), cache.weight(name + "w2", w2)
        scale = float((d // H) ** -0.5)

        u1 = _empty((M, d), BF16, dev)
        stats = _empty((4, M), F32, dev)
        L.layernorm_fwd(x, ln1w.detach(), ln1b.detach(), eps, M, d, u1, stats[0], stats[1])
        qkv = _empty((M, 3 * d), BF16, dev)
        L.gemm(u1, wqkv, M, 3 * d, d, out_bf16=qkv, bias=bqkv)
        attn = _empty((M, d), BF16, dev)
        lse = _empty((B, H, S), F32, dev)
        L.attn_fwd(qkv, B, S, H, scale, attn, lse)
        x_mid = _empty((M, d), F32, dev)
        L.gemm(attn, wob, M, d, d, out_f32=x_mid, bias=bo.detach(), res=x, ldr=d)
        u2 = _empty((M, d), BF16, dev)
        L.layernorm_fwd(x_mid, ln2w.detach(), ln2b.detach(), eps, M, d, u2, stats[2], stats[3])
        pre = _empty((M, ff), BF16, dev)
        act = _empty((M, ff), BF16, dev)
        L.gemm(u2, w1b, M, ff, d, out_bf16=act, bias=b1.detach(), act=1, aux_out=pre, ld_aux=ff)
        x_out = _empty((M, d), F32, dev)
        L.gemm(act, w2b, M, d, ff, out_f32=x_out, bias=b2.detach(), res=x_mid, ldr=d)

        ctx.st, ctx.name, ctx.geom = st, name, (M, d, ff, B, S, H, scale)
        ctx.saved = (x, u1, stats, qkv, attn, lse, x_mid, u2, pre, act, ln1w, ln2w, wqkv, wob, w1b, w2b)
        return x_out

    @staticmethod
    def backward(ctx, dxo):
        st = ctx.st
        M, d, ff, B, S, H, scale = ctx.geom
        x, u1, stats, qkv, attn, lse, x_mid, u2, pre, act, ln1w, ln2w, wqkv, wob, w1b, w2b = ctx.saved
        ctx.saved = None
        dev = dxo.device
        dxo = _contig_grad(dxo)
        dxob = st.side.take(dxo, M, d)

        # every atomically-accumulated output of this block in one zero-filled buffer (one memset)
        sizes = [d, d, 3 * d * d, 3 * d, d * d, d, d, d, ff * d, ff, d * ff, d]
        flat = torch.zeros(sum(sizes), dtype=F32, device=dev)
        views, o = [], 0
        for s_ in sizes:
            views.append(flat[o:o + s_])
            o += s_
        g_ln1w, g_ln1b, g_wqkv, g_bqkv, g_wo, g_bo, g_ln2w, g_ln2b, g_w1, g_b1, g_w2, g_b2 = views

        # fc2: x_out = x_mid + act.W2^T + b2
        d_pre = _empty((M, ff), BF16, dev)
        L.gemm(dxob, w2b, M, ff, d, b_mn=True, ldb=ff, out_bf16=d_pre, act=2, aux_in=pre, ld_aux=ff)
        L.gemm(dxob, act, d, ff, M, a_mn=True, b_mn=True, lda=d, ldb=ff, out_f32=g_w2, k_splits=0)
        L.colsum(dxob, M, d, g_b2)
        del act, pre
        # fc1: pre = u2.W1^T + b1
        L.gemm(d_pre, u2, ff, d, M, a_mn=True, b_mn=True, lda=ff, ldb=d, out_f32=g_w1, k_splits=0)
        L.colsum(d_pre, M, ff, g_b1)
        d_u2 = _empty((M, d), BF16, dev)
        L.gemm(d_pre, w1b, M, d, ff, b_mn=True, ldb=d, out_bf16=d_u2)
        del d_pre, u2
        # LN2 backward + residual branch
        dxm = _empty((M, d), F32, dev)
        dxmb = _empty((M, d), BF16, dev)
        L.layernorm_bwd(d_u2, x_mid, stats[2], stats[3], ln2w.detach(), dxo, M, d, dxm, dxmb, g_ln2w, g_ln2b)
        del d_u2, x_mid
        # attention output projection: x_mid = x + attn.Wo^T + bo
        d_attn = _empty((M, d), BF16, dev)
        L.gemm(dxmb, wob, M, d, d, b_mn=True, ldb=d, out_bf16=d_attn)
        L.gemm(dxmb, attn, d, d, M, a_mn=True, b_mn=True, lda=d, ldb=d, out_f32=g_wo, k_splits=0)
        L.colsum(dxmb, M, d, g_bo)
        # attention core
        dqkv = _empty((M, 3 * d), BF16, dev)
        delta = _empty((B, H, S), F32, dev)
        L.attn_bwd(qkv, attn, d_attn, lse, B, S, H, scale, delta, dqkv)
        del d_attn, attn, qkv
        # fused QKV projection
        L.gemm(dqkv, u1, 3 * d, d, M, a_mn=True, b_mn=True, lda=3 * d, ldb=d, out_f32=g_wqkv, k_splits=0)
        L.colsum(dqkv, M, 3 * d, g_bqkv)
        d_u1 = _empty((M, d), BF16, dev)
        L.gemm(dqkv, wqkv, M, d, 3 * d, b_mn=True, ldb=d, out_bf16=d_u1)
        del dqkv, u1
        # LN1 backward + residual
        dx = _empty((M, d), F32, dev)
        dxb = _empty((M, d), BF16, dev)
        L.layernorm_bwd(d_u1, x, stats[0], stats[1], ln1w.detach(), dxm, M, d, dx, dxb, g_ln1w, g_ln1b)
        st.side.put(dx, dxb)

        gw = g_wqkv.view(3, d, d)
        return (dx, g_ln1w, g_ln1b, gw[0], gw[1], gw[2], g_bqkv[0:d], g_bqkv[2 * d:3 * d], g_wo.view(d, d), g_bo,
                g_ln2w, g_ln2b, g_w1.view(ff, d), g_b1, g_w2.view(d, ff), g_b2, None, None, None, None, None, None)


# ==================================================================================================== enc -> dec
class EncToDecFn(torch.autograd.Function):
    """x_full = cat([W_e2d.h + pos[vis], mask_token + pos[masked]], 1)   (HF:576, HF:585-591), fp32 [B*N, Dd]."""

    @staticmethod
    def forward(ctx, h, w, mask_token, pos, st: StepState, name):
        dev = h.device
        M, D = h.shape
        Dd = w.shape[0]
        B, N, nv = st.B, st.N, st.nv
        wb = st.cache.weight(name, w)
        hb = _empty((M, D), BF16, dev)
        L.rows_to_bf16(h, M, D, hb)
        xf = _empty((B * N, Dd), F32, dev)
        L.gemm(hb, wb, M, Dd, D, out_f32=xf, res=pos, ldr=Dd, res_idx=st.vis_idx, out_seg=nv, out_seg_stride=N,
               out_seg_off=0)
        L.decoder_mask_rows(xf, mask_token.detach().reshape(-1), pos, st.msk_idx, B, N, nv, Dd)
        ctx.st, ctx.hb, ctx.wb, ctx.dims = st, hb, wb, (M, D, Dd)
        return xf

    @staticmethod
    def backward(ctx, dxf):
        st, hb, wb = ctx.st, ctx.hb, ctx.wb
        M, D, Dd = ctx.dims
        B, N, nv, nm = st.B, st.N, st.nv, st.nm
        dev = dxf.device
        dxf = _contig_grad(dxf)
        st.side.drop()
        dz = _empty((M, Dd), BF16, dev)
        L.rows_to_bf16(dxf, M, Dd, dz, ld=Dd, seg=(nv, N, 0))
        flat = torch.zeros(Dd * D + Dd, dtype=F32, device=dev)
        g_w, g_tok = flat[:Dd * D].view(Dd, D), flat[Dd * D:]
        dh = _empty((M, D), F32, dev)
        dhb = _empty((M, D), BF16, dev)
        L.gemm(dz, wb, M, D, Dd, b_mn=True, ldb=D, out_f32=dh, out_bf16=dhb)
        st.side.put(dh, dhb)
        L.gemm(dz, hb, Dd, D, M, a_mn=True, b_mn=True, lda=Dd, ldb=D, out_f32=g_w, k_splits=0)
        L.colsum(dxf, B * nm, Dd, g_tok, ld=Dd, seg=(nm, N, nv))
        return dh, g_w, g_tok.view(1, 1, Dd), None, None, None


# ==================================================================================================== head + loss
class HeadLossFn(torch.autograd.Function):
    """logits = W_h.LN(x[:, -Nm:]) + b_h (HF:506-510); loss = mean((logits - target)^2) (HF:672-673), fused:
    the GEMM epilogue subtracts the target tile, accumulates the squared error and keeps (logits - target) in bf16 for
    the backward, whose scale 2/numel * grad_output is applied inside the dgrad / wgrad / bias-sum kernels."""

    @staticmethod
    def forward(ctx, xf, nw, nb, wh, bh, target, st: StepState, name, want_logits):
        dev = xf.device
        B, N, nv, nm = st.B, st.N, st.nv, st.nm
        Dd = xf.shape[1]
        K = wh.shape[0]
        M = B * nm
        whb = st.cache.weight(name, wh)
        z = _empty((M, Dd), BF16, dev)
        stats = _empty((2, M), F32, dev)
        L.layernorm_fwd(xf, nw.detach(), nb.detach(), 1e-5, M, Dd, z, stats[0], stats[1], ldx=Dd, seg=(nm, N, nv))
        diff = _empty((M, K), BF16, dev)
        logits = _empty((M, K), BF16, dev) if want_logits else None
        part = _empty((L.gemm_loss_slots(M, K, 0),), F32, dev)
        L.gemm(z, whb, M, K, Dd, out_bf16=diff, bias=bh.detach(), target=target, ldt=K, loss_partial=part,
               logits_out=logits)
        loss = _empty((), F32, dev)
        L.loss_finalize(part, float(M) * K, st.status, loss)
        ctx.st, ctx.saved, ctx.dims = st, (xf, stats, z, diff, nw, whb), (M, Dd, K)
        st.logits = logits
        return loss

    @staticmethod
    def backward(ctx, g):
        st = ctx.st
        xf, stats, z, diff, nw, whb = ctx.saved
        ctx.saved = None
        M, Dd, K = ctx.dims
        B, N, nv, nm = st.B, st.N, st.nv, st.nm
        dev = xf.device
        g = g.detach().to(F32).reshape(1).contiguous()
        alpha = 2.0 / (float(M) * K)
        flat = torch.zeros(K * Dd + K + 2 * Dd, dtype=F32, device=dev)
        g_wh, g_bh = flat[:K * Dd].view(K, Dd), flat[K * Dd:K * Dd + K]
        g_nw, g_nb = flat[K * Dd + K:K * Dd + K + Dd], flat[K * Dd + K + Dd:]
        dz = _empty((M, Dd), BF16, dev)
        L.gemm(diff, whb, M, Dd, K, b_mn=True, ldb=Dd, out_bf16=dz, alpha=alpha, alpha_dev=g)
        L.gemm(diff, z, K, Dd, M, a_mn=True, b_mn=True, lda=K, ldb=Dd, out_f32=g_wh, k_splits=0, alpha=alpha,
               alpha_dev=g)
        L.colsum(diff, M, K, g_bh, scale=alpha, scale_dev=g)
        dxf = torch.zeros((B * N, Dd), dtype=F32, device=dev)      # visible rows get no gradient from the head
        dxfb = torch.zeros((B * N, Dd), dtype=BF16, device=dev)
        L.layernorm_bwd(dz, xf, stats[0], stats[1], nw.detach(), None, M, Dd, dxf, dxfb, g_nw, g_nb, ldx=Dd,
                        seg=(nm, N, nv))
        st.side.put(dxf, dxfb)
        return dxf, g_nw, g_nb, g_wh, g_bh, None, None, None, None
