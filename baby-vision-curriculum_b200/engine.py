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

import os

import torch

from . import _lib as L

BF16 = torch.bfloat16
F32 = torch.float32


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


class Bf16Cache:
    """bf16 copies of the fp32 master weights (what autocast re-casts on every call, HF linear layers) in ONE flat
    buffer, refreshed by a single bvc_cast_multi launch whenever any parameter's storage or version counter changed
    (i.e. after an optimizer step / load_state_dict).  Also packs (q_bias, 0, v_bias) into the fused QKV bias."""

    def __init__(self):
        self._plan = None       # (ptr signature, table tensor, n_entries, total elems, views)
        self._versions = None
        self._params = []
        self.epoch = 0          # bumped whenever the copies are rewritten (refresh / optimizer shadow update)

    def register(self, entries, device):
        """entries: list of (name, kind, params) with kind 'w' (one weight) or 'qkv' (wq, wk, wv, qb, vb)."""
        sig = tuple(p.data_ptr() for _, _, ps in entries for p in ps) + (str(device),)
        if self._plan is not None and self._plan[0] == sig:
            return
        n_bf16 = 0
        n_f32 = 0
        for _, kind, ps in entries:
            if kind == "w":
                n_bf16 += (ps[0].numel() + 7) // 8 * 8
            else:
                d = ps[0].shape[0]
                n_bf16 += 3 * d * d
                n_f32 += 3 * d
        wbuf = torch.empty(n_bf16, dtype=BF16, device=device)
        bbuf = torch.zeros(max(n_f32, 8), dtype=F32, device=device)
        rows, views, ow, ob, total = [], {}, 0, 0, 0
        self._params = []

        def add(src, dst_ptr, n, is_f32):
            nonlocal total
            rows.append((src.data_ptr(), dst_ptr, n, 1 if is_f32 else 0))
            total += n
            self._params.append(src)

        for name, kind, ps in entries:
            if kind == "w":
                n = ps[0].numel()
                views[name] = wbuf[ow:ow + n]
                add(ps[0], wbuf.data_ptr() + 2 * ow, n, False)
                ow += (n + 7) // 8 * 8
            else:
                wq, wk, wv, qb, vb = ps
                d = wq.shape[0]
                views[name] = (wbuf[ow:ow + 3 * d * d], bbuf[ob:ob + 3 * d])
                for i, w in enumerate((wq, wk, wv)):
                    add(w, wbuf.data_ptr() + 2 * (ow + i * d * d), d * d, False)
                add(qb, bbuf.data_ptr() + 4 * ob, d, True)
                add(vb, bbuf.data_ptr() + 4 * (ob + 2 * d), d, True)
                ow += 3 * d * d
                ob += 3 * d
        import numpy as np
        tab = np.zeros((len(rows), 4), dtype=np.int64)
        for i, (sp, dp, n, f) in enumerate(rows):
            tab[i] = (sp, dp, n, f)  # {src, dst, n, dst_is_f32 | pad<<32} : 4 x 8 bytes
        table = torch.from_numpy(tab).to(device)
        self._plan = (sig, table, len(rows), total, views, wbuf, bbuf)
        self._versions = None
        self._shadow = {id(src): (dp, f) for src, (_, dp, _, f) in zip(self._params, rows)}

    def refresh(self):
        ver = tuple(p._version for p in self._params)
        if ver != self._versions:
            _, table, n, total, _, _, _ = self._plan
            L.cast_multi(table, n, total)
            self._versions = ver
            self.epoch += 1

    def shadows(self):
        """{id(param): (copy pointer, copy_is_f32)} when every copy matches its parameter right now, else {} --
        FusedSGD (optim.py) updates the copies in the same pass as the parameters."""
        if self._plan is None or self._versions is None:
            return {}
        if tuple(p._version for p in self._params) != self._versions:
            return {}
        return self._shadow

    def mark_synced(self):
        self._versions = tuple(p._version for p in self._params)
        self.epoch += 1

    def weight(self, name):
        return self._plan[4][name]

    def qkv(self, name):
        return self._plan[4][name]

    def clear(self):
        self._plan, self._versions, self._params = None, None, []


class GradSideChannel:
    """Companions of the most recent fp32 activation gradient, written by the kernel that produced it: its bf16 twin
    (the next GEMM's A operand) and, when available, its column sums (the bias gradient of the Linear that fed this
    residual stream) -- so the upstream stage neither re-casts nor re-reads its incoming gradient."""

    def __init__(self):
        self.ptr = None
        self.t = None
        self.colsum = None

    def put(self, g_f32, g_bf16, colsum=None):
        self.ptr, self.t, self.colsum = g_f32.data_ptr(), g_bf16, colsum

    def drop(self):
        self.ptr, self.t, self.colsum = None, None, None

    def take(self, g_f32, M, d):
        """-> (bf16 twin, column sums or None)"""
        if self.ptr == g_f32.data_ptr() and self.t is not None and self.t.numel() == M * d:
            t, cs = self.t, self.colsum
            self.drop()
            return t, cs
        self.drop()
        out = _empty((M, d), BF16, g_f32.device)
        L.rows_to_bf16(g_f32, M, d, out)
        return out, None


class StepState:
    """Per-forward shared state (index lists, shapes, caches)."""

    def __init__(self, cache: Bf16Cache):
        self.cache = cache
        self.side = GradSideChannel()
        self.B = self.N = self.nv = self.nm = 0
        self.vis_idx = self.msk_idx = self.slot = self.status = None
        self.sync = None  # ddp.GradSync when the model is wrapped in bvc_b200.DistributedDataParallel
        self.epoch = cache.epoch  # the stages' backward uses the SAME bf16 weight copies the forward read (see check())
        self.stage_params = {}  # stage name -> its parameters, in the order the stage's backward returns their gradients

    def check(self):
        """The stages keep raw references to the shared bf16 weight copies instead of save_for_backward; an optimizer
        step / refresh between a forward and its backward rewrites them in place.  torch raises a version-counter
        error in that situation -- so do we."""
        if self.cache.epoch != self.epoch:
            raise RuntimeError("one of the variables needed for gradient computation has been modified by an inplace "
                               "operation: the model's weights (bf16 operand copies) were updated between this forward "
                               "and its backward")

    def reduce(self, flat, name=None, views=None):
        """A stage's parameter gradients (one contiguous buffer) are complete on the stream: start their all-reduce.
        `views` are the gradient tensors the stage returns to autograd, in the order of stage_params[name]."""
        if self.sync is not None:
            self.sync.reduce(flat, self.stage_params.get(name), views)


def _contig_grad(g):
    return g if g.is_contiguous() else g.contiguous()


# Weight-gradient GEMMs of a block do not feed the rest of its backward.  Launched on a second stream, each one starts
# on the SMs that the main stream's current GEMM leaves idle in its last, partial wave of tiles (the kernels are
# persistent with one CTA per SM, and e.g. the 240 tiles of an M = 10240, N = 768 GEMM fill 148 SMs to 81 %), instead of
# queueing behind it.  BVC_WGRAD_STREAM=0 keeps everything on one stream.
_WGRAD_STREAM = os.environ.get("BVC_WGRAD_STREAM", "1") != "0"
_wgrad_streams = {}


def set_wgrad_stream(on: bool) -> bool:
    """Enable / disable the second stream (returns the previous setting).  bench.py turns it off for its per-kernel
    timing pass: concurrent kernels share the SMs, so their individual event durations stop being kernel times."""
    global _WGRAD_STREAM
    old, _WGRAD_STREAM = _WGRAD_STREAM, bool(on)
    return old


class _SideStream:
    """`with side.after_main(): launch(...)` enqueues on the side stream behind everything queued so far on the
    current stream; join() makes the current stream wait for the side stream."""

    def __init__(self, dev):
        self.stream = _wgrad_streams.get(dev)
        if self.stream is None:
            self.stream = _wgrad_streams[dev] = torch.cuda.Stream(device=dev)
        self.used = False

    def after_main(self):
        ev = torch.cuda.Event()
        ev.record()
        self.stream.wait_event(ev)
        self.used = True
        return torch.cuda.stream(self.stream)

    def join(self):
        if self.used:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.used = False


class _NoSideStream:
    def after_main(self):
        import contextlib
        return contextlib.nullcontext()

    def join(self):
        pass


# ==================================================================================================== embed
class EmbedFn(torch.autograd.Function):
    """e = W_pe . patch + b_pe + pos[vis_idx]  on the visible tubelets only (HF:164-177, HF:109-122)."""

    @staticmethod
    def forward(ctx, w, b, patches, pos, st: StepState, name):
        D = w.shape[0]
        K = w.numel() // D
        M = patches.shape[0]
        wb = st.cache.weight(name)
        x = _empty((M, D), F32, patches.device)
        L.gemm(patches, wb, M, D, K, out_f32=x, bias=b.detach(), res=pos, ldr=D, res_idx=st.vis_idx)
        ctx.st, ctx.patches, ctx.dims, ctx.wshape, ctx.name = st, patches, (M, D, K), w.shape, name
        return x

    @staticmethod
    def backward(ctx, dx):
        st, patches = ctx.st, ctx.patches
        st.check()
        M, D, K = ctx.dims
        dx = _contig_grad(dx)
        dxb, cs = st.side.take(dx, M, D)
        flat = torch.zeros(D * K + D, dtype=F32, device=dx.device)
        dw, db = flat[:D * K].view(ctx.wshape), flat[D * K:]
        L.gemm(dxb, patches, D, K, M, a_mn=True, b_mn=True, lda=D, ldb=K, out_f32=dw, k_splits=0)
        if cs is not None:
            db = cs
        else:
            L.colsum(dxb, M, D, db)
        st.reduce(flat, ctx.name, (dw, db))
        return dw, db, None, None, None, None


# ==================================================================================================== block
def _block_fwd(ctx, x, ln1w, ln1b, wqkv, bqkv, wob, bo, ln2w, ln2b, w1b, b1, w2b, b2, st, name, B, S, H, eps):
    """One pre-LN transformer block on bf16 operand copies wqkv [3d, d] / wob / w1b / w2b and fp32 biases:
    x + Wo.Attn(LN1(x)) then + W2.gelu(W1.LN2(.)).  Shared by the HF-layout block (separate q / k / v weights, no k
    bias: BlockFn) and the fused-qkv block of the predictive path (FusedQkvBlockFn)."""
    dev = x.device
    M, d = x.shape
    ff = w1b.numel() // d
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
    gp = _empty((M, ff), BF16, dev)   # gelu'(pre-activation): all the backward needs of it
    act = _empty((M, ff), BF16, dev)
    L.gemm(u2, w1b, M, ff, d, out_bf16=act, bias=b1.detach(), act=1, aux_out=gp, ld_aux=ff)
    x_out = _empty((M, d), F32, dev)
    L.gemm(act, w2b, M, d, ff, out_f32=x_out, bias=b2.detach(), res=x_mid, ldr=d)

    ctx.st, ctx.name, ctx.geom = st, name, (M, d, ff, B, S, H, scale)
    ctx.saved = (x, u1, stats, qkv, attn, lse, x_mid, u2, gp, act, ln1w, ln2w, wqkv, wob, w1b, w2b)
    return x_out


def _block_bwd(ctx, dxo):
    """-> dx, flat gradient buffer, (g_ln1w, g_ln1b, g_wqkv [3 d d], g_bqkv [3 d], g_wo [d d], g_bo, g_ln2w, g_ln2b,
    g_w1 [ff d], g_b1, g_w2 [d ff], g_b2): views of `flat` (g_b2 may live in the downstream stage's buffer)."""
    st = ctx.st
    st.check()
    M, d, ff, B, S, H, scale = ctx.geom
    x, u1, stats, qkv, attn, lse, x_mid, u2, gp, act, ln1w, ln2w, wqkv, wob, w1b, w2b = ctx.saved
    ctx.saved = None
    dev = dxo.device
    dxo = _contig_grad(dxo)
    dxob, cs_out = st.side.take(dxo, M, d)
    wg = _SideStream(dev) if _WGRAD_STREAM else _NoSideStream()  # operands stay referenced until wg.join() below

    # every atomically-accumulated output of this block in one zero-filled buffer (= the all-reduce bucket).  The
    # vectors (accumulated by main-stream kernels; the last slot collects colsum(dx) of this block's input gradient for the
    # stage upstream) come first and are cleared here; the four weight matrices -- 99.9 % of the bytes, written only by
    # the weight-gradient GEMMs on the side stream -- are cleared there, off the critical path.
    sizes = [d, d, 3 * d, d, d, d, ff, d, d, 3 * d * d, d * d, ff * d, d * ff]
    n_vec = sum(sizes[:9])
    flat = torch.empty(sum(sizes), dtype=F32, device=dev)
    flat[:n_vec].zero_()
    with wg.after_main():
        flat[n_vec:].zero_()
    views, o = [], 0
    for s_ in sizes:
        views.append(flat[o:o + s_])
        o += s_
    g_ln1w, g_ln1b, g_bqkv, g_bo, g_ln2w, g_ln2b, g_b1, g_b2, cs_in, g_wqkv, g_wo, g_w1, g_w2 = views

    # long sequences (the decoder's 1568 tokens): one-pass attention backward, dQ contributions reduced into an fp32
    # scratch.  The reduction order of those fp32 adds (TMA reduce, L2 atomics) varies from run to run -- like torch's own
    # flash attention backward -- so two backward passes agree to ~2e-5 (global rel-L2; up to ~4e-4 on a single deep
    # tensor, bf16 roundings downstream amplify last-bit differences) instead of bit for bit.  Under
    # torch.use_deterministic_algorithms(True) (or BVC_ATTN_BWD1=0) the deterministic two-pass kernels run instead.
    # The scratch is cleared NOW, on the side stream, under the MLP's backward GEMMs (154 MB of writes off the
    # critical path: 40 us per decoder layer).
    one_pass = S > 160 and not torch.are_deterministic_algorithms_enabled()
    dq_acc, dq_zeroed = None, None
    if one_pass:
        dq_acc = _empty((B, S, H, 64), F32, dev)
        if isinstance(wg, _SideStream):
            with wg.after_main():
                dq_acc.zero_()
                dq_zeroed = torch.cuda.Event()
                dq_zeroed.record()

    # fc2: x_out = x_mid + act.W2^T + b2
    d_pre = _empty((M, ff), BF16, dev)
    # x gelu' (saved by the forward epilogue) fused; the epilogue also accumulates the column sums of d_pre = the
    # fc1 bias gradient
    L.gemm(dxob, w2b, M, ff, d, b_mn=True, ldb=ff, out_bf16=d_pre, act=2, aux_in=gp, ld_aux=ff, colsum=g_b1)
    with wg.after_main():
        L.gemm(dxob, act, d, ff, M, a_mn=True, b_mn=True, lda=d, ldb=ff, out_f32=g_w2, k_splits=0)
    if cs_out is not None:
        g_b2 = cs_out
    else:
        L.colsum(dxob, M, d, g_b2)
    # fc1: pre = u2.W1^T + b1
    d_u2 = _empty((M, d), BF16, dev)
    L.gemm(d_pre, w1b, M, d, ff, b_mn=True, ldb=d, out_bf16=d_u2)
    with wg.after_main():
        L.gemm(d_pre, u2, ff, d, M, a_mn=True, b_mn=True, lda=ff, ldb=d, out_f32=g_w1, k_splits=0)
    # LN2 backward + residual branch
    dxm = _empty((M, d), F32, dev)
    dxmb = _empty((M, d), BF16, dev)
    L.layernorm_bwd(d_u2, x_mid, stats[2], stats[3], ln2w.detach(), dxo, M, d, dxm, dxmb, g_ln2w, g_ln2b,
                    dxsum=g_bo)  # colsum(dx_mid) is the out-proj bias gradient
    del d_u2, x_mid
    # attention output projection: x_mid = x + attn.Wo^T + bo
    d_attn = _empty((M, d), BF16, dev)
    L.gemm(dxmb, wob, M, d, d, b_mn=True, ldb=d, out_bf16=d_attn)
    with wg.after_main():
        L.gemm(dxmb, attn, d, d, M, a_mn=True, b_mn=True, lda=d, ldb=d, out_f32=g_wo, k_splits=0)
    # attention core
    dqkv = _empty((M, 3 * d), BF16, dev)
    delta = _empty((B, H, S), F32, dev)
    if dq_zeroed is not None:
        torch.cuda.current_stream().wait_event(dq_zeroed)
    L.attn_bwd(qkv, attn, d_attn, lse, B, S, H, scale, delta, dqkv, dq_acc, dq_accum_zeroed=dq_zeroed is not None)
    del dq_acc
    # fused QKV projection
    d_u1 = _empty((M, d), BF16, dev)
    L.gemm(dqkv, wqkv, M, d, 3 * d, b_mn=True, ldb=d, out_bf16=d_u1)
    with wg.after_main():
        L.gemm(dqkv, u1, 3 * d, d, M, a_mn=True, b_mn=True, lda=3 * d, ldb=d, out_f32=g_wqkv, k_splits=0)
        L.colsum(dqkv, M, 3 * d, g_bqkv)
    # LN1 backward + residual
    dx = _empty((M, d), F32, dev)
    dxb = _empty((M, d), BF16, dev)
    L.layernorm_bwd(d_u1, x, stats[0], stats[1], ln1w.detach(), dxm, M, d, dx, dxb, g_ln1w, g_ln1b, dxsum=cs_in)
    st.side.put(dx, dxb, cs_in)
    wg.join()  # the weight gradients are complete on the current stream from here on
    return dx, flat, (g_ln1w, g_ln1b, g_wqkv, g_bqkv, g_wo.view(d, d), g_bo, g_ln2w, g_ln2b, g_w1.view(ff, d), g_b1,
                      g_w2.view(d, ff), g_b2)


class BlockFn(torch.autograd.Function):
    """One pre-LN transformer block (HF:348-366) in HF's parameter layout: separate query / key / value weights, q and
    v biases (the key has none, HF:222-224)."""

    @staticmethod
    def forward(ctx, x, ln1w, ln1b, wq, wk, wv, qb, vb, wo, bo, ln2w, ln2b, w1, b1, w2, b2, st: StepState, name, B, S,
                H, eps):
        cache = st.cache
        wqkv, bqkv = cache.qkv(name + "qkv")
        wob, w1b, w2b = cache.weight(name + "wo"), cache.weight(name + "w1"), cache.weight(name + "w2")
        return _block_fwd(ctx, x, ln1w, ln1b, wqkv, bqkv, wob, bo, ln2w, ln2b, w1b, b1, w2b, b2, st, name, B, S, H, eps)

    @staticmethod
    def backward(ctx, dxo):
        d = ctx.geom[1]
        dx, flat, (g_ln1w, g_ln1b, g_wqkv, g_bqkv, g_wo, g_bo, g_ln2w, g_ln2b, g_w1, g_b1, g_w2, g_b2) = \
            _block_bwd(ctx, dxo)
        gw = g_wqkv.view(3, d, d)
        grads = (g_ln1w, g_ln1b, gw[0], gw[1], gw[2], g_bqkv[0:d], g_bqkv[2 * d:3 * d], g_wo, g_bo, g_ln2w, g_ln2b, g_w1,
                 g_b1, g_w2, g_b2)
        ctx.st.reduce(flat, ctx.name, grads)
        return (dx,) + grads + (None, None, None, None, None, None)


class FusedQkvBlockFn(torch.autograd.Function):
    """The same block in the predictive path's layout (pretraining/predictive/vision_transformer.py:186-231): ONE
    `qkv = nn.Linear(dim, 3 dim, bias=qkv_bias)` whose output reshapes to [B, N, 3, heads, head_dim] -- exactly the
    fused-QKV layout of the attention kernels -- LayerNorm eps 1e-6, MLP fc1 / GELU / fc2.  bqkv may be None."""

    @staticmethod
    def forward(ctx, x, ln1w, ln1b, wqkv, bqkv, wo, bo, ln2w, ln2b, w1, b1, w2, b2, st: StepState, name, B, S, H, eps):
        cache = st.cache
        ctx.has_qkv_bias = bqkv is not None
        return _block_fwd(ctx, x, ln1w, ln1b, cache.weight(name + "qkv"), bqkv.detach() if bqkv is not None else None,
                          cache.weight(name + "wo"), bo, ln2w, ln2b, cache.weight(name + "w1"), b1,
                          cache.weight(name + "w2"), b2, st, name, B, S, H, eps)

    @staticmethod
    def backward(ctx, dxo):
        d = ctx.geom[1]
        dx, flat, (g_ln1w, g_ln1b, g_wqkv, g_bqkv, g_wo, g_bo, g_ln2w, g_ln2b, g_w1, g_b1, g_w2, g_b2) = \
            _block_bwd(ctx, dxo)
        grads = (g_ln1w, g_ln1b, g_wqkv.view(3 * d, d), g_bqkv if ctx.has_qkv_bias else None, g_wo, g_bo, g_ln2w, g_ln2b,
                 g_w1, g_b1, g_w2, g_b2)
        ctx.st.reduce(flat, ctx.name, grads)
        return (dx,) + grads + (None, None, None, None, None, None)


# ==================================================================================================== enc -> dec
class EncToDecFn(torch.autograd.Function):
    """x_full = cat([W_e2d.h + pos[vis], mask_token + pos[masked]], 1)   (HF:576, HF:585-591), fp32 [B*N, Dd]."""

    @staticmethod
    def forward(ctx, h, w, mask_token, pos, st: StepState, name):
        dev = h.device
        M, D = h.shape
        Dd = w.shape[0]
        B, N, nv = st.B, st.N, st.nv
        wb = st.cache.weight(name)
        hb = _empty((M, D), BF16, dev)
        L.rows_to_bf16(h, M, D, hb)
        xf = _empty((B * N, Dd), F32, dev)
        L.gemm(hb, wb, M, Dd, D, out_f32=xf, res=pos, ldr=Dd, res_idx=st.vis_idx, out_seg=nv, out_seg_stride=N,
               out_seg_off=0)
        L.decoder_mask_rows(xf, mask_token.detach().reshape(-1), pos, st.msk_idx, B, N, nv, Dd)
        ctx.st, ctx.hb, ctx.wb, ctx.dims, ctx.name = st, hb, wb, (M, D, Dd), name
        return xf

    @staticmethod
    def backward(ctx, dxf):
        st, hb, wb = ctx.st, ctx.hb, ctx.wb
        st.check()
        M, D, Dd = ctx.dims
        B, N, nv, nm = st.B, st.N, st.nv, st.nm
        dev = dxf.device
        dxf = _contig_grad(dxf)
        st.side.drop()
        dz = _empty((M, Dd), BF16, dev)
        L.rows_to_bf16(dxf, M, Dd, dz, ld=Dd, seg=(nv, N, 0))
        flat = torch.zeros(Dd * D + Dd, dtype=F32, device=dev)
        g_w, g_tok = flat[:Dd * D].view(Dd, D), flat[Dd * D:Dd * D + Dd]
        dh = _empty((M, D), F32, dev)
        dhb = _empty((M, D), BF16, dev)
        L.gemm(dz, wb, M, D, Dd, b_mn=True, ldb=D, out_f32=dh, out_bf16=dhb)
        st.side.put(dh, dhb)
        L.gemm(dz, hb, Dd, D, M, a_mn=True, b_mn=True, lda=Dd, ldb=D, out_f32=g_w, k_splits=0)
        L.colsum(dxf, B * nm, Dd, g_tok, ld=Dd, seg=(nm, N, nv))
        g_tok = g_tok.view(1, 1, Dd)
        st.reduce(flat, ctx.name, (g_w, g_tok))
        return dh, g_w, g_tok, None, None, None


# ==================================================================================================== head + loss
class HeadLossFn(torch.autograd.Function):
    """logits = W_h.LN(x[:, -Nm:]) + b_h (HF:506-510); loss = mean((logits - target)^2) (HF:672-673), fused:
    the GEMM epilogue subtracts the target tile, accumulates the squared error and keeps (logits - target) in bf16 for
    the backward, whose scale 2/numel * grad_output is applied inside the dgrad / wgrad / bias-sum kernels."""

    @staticmethod
    def forward(ctx, xf, nw, nb, wh, bh, target, st: StepState, name):
        dev = xf.device
        B, N, nv, nm = st.B, st.N, st.nv, st.nm
        Dd = xf.shape[1]
        K = wh.shape[0]
        M = B * nm
        whb = st.cache.weight(name)
        z = _empty((M, Dd), BF16, dev)
        stats = _empty((2, M), F32, dev)
        L.layernorm_fwd(xf, nw.detach(), nb.detach(), 1e-5, M, Dd, z, stats[0], stats[1], ldx=Dd, seg=(nm, N, nv))
        diff = _empty((M, K), BF16, dev)
        bn = 192 if K % 192 == 0 else 0  # measured best for the short-K head GEMM (profiles/r01_gemm_tune.log)
        part = _empty((L.gemm_loss_slots(M, K, bn),), F32, dev)
        # the logits themselves are not written (277 MB per step at batch 64 that the training loop never reads): the
        # output object re-runs the projection on z if `.logits` is accessed (modeling_videomae.py)
        L.gemm(z, whb, M, K, Dd, out_bf16=diff, bias=bh.detach(), target=target, ldt=K, loss_partial=part,
               block_n=bn)
        loss = _empty((), F32, dev)
        L.loss_finalize(part, float(M) * K, st.status, loss)
        ctx.st, ctx.saved, ctx.dims, ctx.name = st, (xf, stats, z, diff, nw, whb), (M, Dd, K), name
        st.head_operands = (z, whb, bh, bn)
        return loss

    @staticmethod
    def backward(ctx, g):
        st = ctx.st
        st.check()
        xf, stats, z, diff, nw, whb = ctx.saved
        ctx.saved = None
        M, Dd, K = ctx.dims
        B, N, nv, nm = st.B, st.N, st.nv, st.nm
        dev = xf.device
        g = g.detach().to(F32).reshape(1).contiguous()
        alpha = 2.0 / (float(M) * K)
        flat = torch.zeros(K * Dd + K + 3 * Dd, dtype=F32, device=dev)
        g_wh, g_bh = flat[:K * Dd].view(K, Dd), flat[K * Dd:K * Dd + K]
        o = K * Dd + K
        g_nw, g_nb, cs = flat[o:o + Dd], flat[o + Dd:o + 2 * Dd], flat[o + 2 * Dd:]
        dz = _empty((M, Dd), BF16, dev)
        L.gemm(diff, whb, M, Dd, K, b_mn=True, ldb=Dd, out_bf16=dz, alpha=alpha, alpha_dev=g)
        L.gemm(diff, z, K, Dd, M, a_mn=True, b_mn=True, lda=K, ldb=Dd, out_f32=g_wh, k_splits=0, alpha=alpha,
               alpha_dev=g)
        L.colsum(diff, M, K, g_bh, scale=alpha, scale_dev=g)
        dxf = torch.zeros((B * N, Dd), dtype=F32, device=dev)      # visible rows get no gradient from the head
        dxfb = torch.zeros((B * N, Dd), dtype=BF16, device=dev)
        L.layernorm_bwd(dz, xf, stats[0], stats[1], nw.detach(), None, M, Dd, dxf, dxfb, g_nw, g_nb, ldx=Dd,
                        seg=(nm, N, nv), dxsum=cs)  # visible rows are zero: colsum over the Nm rows == over all rows
        st.side.put(dxf, dxfb, cs)
        st.reduce(flat, ctx.name, (g_nw, g_nb, g_wh, g_bh))
        return dxf, g_nw, g_nb, g_wh, g_bh, None, None, None
