"""The predictive path's ViT block on the VideoMAE step's kernels.

`Block` is a drop-in for pretraining/predictive/vision_transformer.py:213-231 (`Block`: norm1 -> Attention with ONE
fused `qkv` Linear (:186-210) -> residual -> norm2 -> MLP fc1 / GELU / fc2 (:167-183) -> residual; LayerNorm eps 1e-6 as
configured by vit_base(): `partial(nn.LayerNorm, eps=1e-6)`): same constructor arguments, same sub-module / parameter
names (so the reference's checkpoints load and its optimizers / EMA loop see the same parameters), forward
`[B, N, C] -> [B, N, C]`.  The arithmetic is engine.FusedQkvBlockFn: LayerNorm kernels, tcgen05 GEMMs with fused bias /
GELU / residual epilogues, tcgen05 flash attention (the fused-qkv output reshaped [B, N, 3, heads, 64] IS the attention
kernels' input layout), hand-written backward.  Dtype flow as the VideoMAE engine's: fp32 residual stream, bf16 GEMM /
attention operands, fp32 accumulation and statistics (the reference under autocast: fp32 residual, bf16 Linear outputs).

`convert_blocks(vit)` swaps the blocks of a reference `VisionTransformer` (the context / target encoder of
pretrain_jepa.py:361-433, whose forward :378-402 stays the reference's own) in place, sharing the parameters.

Not covered: head_dim != 64 (the attention kernels are head_dim-64 builds; the reference's predictor runs 384 / 12 =
32-wide heads and keeps torch's blocks), dropout / drop-path > 0 (the reference trains with 0), return_attention.
CUDA only: there is no CPU path.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from .engine import Bf16Cache, FusedQkvBlockFn, GradSideChannel, StepState

_side = {}  # device -> GradSideChannel shared by consecutive blocks (bf16 twin / column sums of the activation gradient)


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the arithmetic runs in libbvc.so through bvc_b200.jepa_vit.Block")


class Attention(_Holder):
    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class MLP(_Holder):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        if dim % num_heads or dim // num_heads != 64:
            raise NotImplementedError("bvc_b200.jepa_vit.Block: the sm_100a attention kernels are built for head_dim 64")
        if qk_scale is not None or drop != 0. or attn_drop != 0. or drop_path != 0.:
            raise NotImplementedError("qk_scale / dropout / drop-path are not on the reference's configuration")
        if act_layer is not nn.GELU:
            raise NotImplementedError("only nn.GELU (exact erf), the reference's activation")
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = MLP(dim, int(dim * mlp_ratio))
        self._cache = Bf16Cache()

    @classmethod
    def from_reference(cls, blk):
        """Adopt the sub-modules (and so the parameters) of a reference `Block` (vision_transformer.py:213-231)."""
        dim = blk.norm1.weight.shape[0]
        heads = blk.attn.num_heads
        if dim // heads != 64:
            raise NotImplementedError("head_dim must be 64")
        if abs(blk.attn.scale - (dim // heads) ** -0.5) > 1e-12:
            raise NotImplementedError("qk_scale override")
        for m in blk.modules():
            if isinstance(m, nn.Dropout) and m.p != 0.:
                raise NotImplementedError("dropout > 0")
        if not isinstance(blk.drop_path, nn.Identity):
            raise NotImplementedError("drop-path > 0")
        if not isinstance(blk.mlp.act, nn.GELU) or getattr(blk.mlp.act, "approximate", "none") != "none":
            raise NotImplementedError("only exact GELU")
        self = cls.__new__(cls)
        nn.Module.__init__(self)
        self.norm1, self.attn, self.drop_path, self.norm2, self.mlp = blk.norm1, blk.attn, blk.drop_path, blk.norm2, blk.mlp
        self._cache = Bf16Cache()
        return self

    def forward(self, x, return_attention=False):
        if return_attention:
            raise NotImplementedError("return_attention: the attention matrix is never materialised")
        if not x.is_cuda:
            raise L.BvcError("bvc_b200.jepa_vit.Block runs on CUDA only; there is no CPU path")
        if x.dim() != 3 or x.shape[2] != self.norm1.weight.shape[0]:
            raise ValueError("x must be [batch, tokens, dim]")
        B, N, C = x.shape
        dev = x.device
        a, m = self.attn, self.mlp
        x2 = x.reshape(B * N, C)
        if x2.dtype != torch.float32:
            x2 = x2.float()
        x2 = x2.contiguous()
        with torch.cuda.device(dev):
            self._cache.register([("qkv", "w", (a.qkv.weight,)), ("wo", "w", (a.proj.weight,)),
                                  ("w1", "w", (m.fc1.weight,)), ("w2", "w", (m.fc2.weight,))], dev)
            self._cache.refresh()
            st = StepState(self._cache)
            st.epoch = self._cache.epoch
            st.side = _side.setdefault(dev, GradSideChannel())
            y = FusedQkvBlockFn.apply(x2, self.norm1.weight, self.norm1.bias, a.qkv.weight, a.qkv.bias, a.proj.weight,
                                      a.proj.bias, self.norm2.weight, self.norm2.bias, m.fc1.weight, m.fc1.bias,
                                      m.fc2.weight, m.fc2.bias, st, "", B, N, a.num_heads, float(self.norm1.eps))
        return y.view(B, N, C)


def convert_blocks(vit):
    """Replace every `Block` of a reference VisionTransformer / VisionTransformerPredictor-like module (attribute
    `blocks`) by a bvc Block sharing its parameters; returns the module."""
    for i, blk in enumerate(vit.blocks):
        if not isinstance(blk, Block):
            vit.blocks[i] = Block.from_reference(blk)
    return vit
