"""Drop-in replacement for transformers.VideoMAEForPreTraining on the reference's hot path.

Boundary (pretraining/generative/pretrain_videomae.py:301):
    outputs = xmodel(inputs, bool_masked_pos=bool_masked_pos);  loss = outputs.loss
Contract kept: constructed from a VideoMAEConfig-like object, `.config` readable,
HF-identical parameter names / shapes / init (state-dicts load both ways, strict), real nn.Parameters so the model
wraps in torch DDP, every parameter receives a gradient each step, loss is an fp32 0-dim tensor with autograd
history that honours an arbitrary upstream grad (GradScaler), ValueError on channel / size mismatch or missing mask.
The module tree below only *holds* parameters under HF's names; all arithmetic is libbvc.so (engine.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _lib as L
from .engine import BF16, F32, Bf16Cache, BlockFn, EmbedFn, EncToDecFn, HeadLossFn, StepState


@dataclass
class VideoMAEConfig:
    """The fields of transformers.VideoMAEConfig this path reads, with HF's defaults
    (transformers/models/videomae/configuration_videomae.py:60-80).  A real transformers.VideoMAEConfig works too."""

    image_size: int = 224
    patch_size: int = 16
    num_channels: int = 3
    num_frames: int = 16
    tubelet_size: int = 2
    hidden_size: int = 768
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = "gelu"
    hidden_dropout_prob: float = 0.0
    attention_probs_dropout_prob: float = 0.0
    initializer_range: float = 0.02
    layer_norm_eps: float = 1e-12
    qkv_bias: bool = True
    use_mean_pooling: bool = True
    decoder_num_attention_heads: int = 6
    decoder_hidden_size: int = 384
    decoder_num_hidden_layers: int = 4
    decoder_intermediate_size: int = 1536
    norm_pix_loss: bool = True


class VideoMAEForPreTrainingOutput:
    """HF:64-75 (`loss`, `logits`, `hidden_states`, `attentions`; tuple-style indexing over the non-None fields).

    `logits` ([B, Nm, 1536] bf16, what HF's autocast forward returns) is materialised on first access: the training
    loop (pretrain_videomae.py:301-302) reads only `.loss`, and the fused head + MSE kernel needs the logits only in
    registers -- writing them every step cost 277 MB of HBM traffic at batch 64 that nobody read.  The first access runs
    the head projection once more on the saved decoder output (same kernel, same tile shape, same accumulation order:
    the values are the ones the loss was computed from); it must happen before the weights are updated."""

    def __init__(self, loss=None, logits=None, hidden_states=None, attentions=None, logits_fn=None):
        self.loss = loss
        self._logits = logits
        self._logits_fn = logits_fn
        self.hidden_states = hidden_states
        self.attentions = attentions

    @property
    def logits(self):
        if self._logits is None and self._logits_fn is not None:
            self._logits = self._logits_fn()
            self._logits_fn = None
        return self._logits

    def __getitem__(self, i):
        return tuple(v for v in (self.loss, self.logits, self.hidden_states, self.attentions) if v is not None)[i]

    def __repr__(self):
        return f"VideoMAEForPreTrainingOutput(loss={self.loss!r}, logits=<{'ready' if self._logits is not None else 'lazy'}>)"


def get_sinusoid_encoding_table(n_position: int, d_hid: int) -> torch.Tensor:
    """HF:78-91 in closed form (float64 through numpy, stored fp32) -- bit-equal to HF's table. [n_position, d_hid]."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)
    ang = pos / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    ang[:, 0::2] = np.sin(ang[:, 0::2])
    ang[:, 1::2] = np.cos(ang[:, 1::2])
    return torch.from_numpy(ang).to(torch.float32)


# ---------------------------------------------------------------------------------------------------------------
# parameter containers under HF's module / parameter names (never called)
# ---------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: the arithmetic runs in libbvc.so through VideoMAEForPreTraining")


class VideoMAEPatchEmbeddings(_Holder):
    def __init__(self, c):
        super().__init__()
        self.projection = nn.Conv3d(c.num_channels, c.hidden_size, kernel_size=(c.tubelet_size, c.patch_size, c.patch_size),
                                    stride=(c.tubelet_size, c.patch_size, c.patch_size))


class VideoMAEEmbeddings(_Holder):
    def __init__(self, c):
        super().__init__()
        self.patch_embeddings = VideoMAEPatchEmbeddings(c)


class VideoMAESelfAttention(_Holder):
    def __init__(self, d):
        super().__init__()
        self.query = nn.Linear(d, d, bias=False)
        self.key = nn.Linear(d, d, bias=False)
        self.value = nn.Linear(d, d, bias=False)
        self.q_bias = nn.Parameter(torch.zeros(d))
        self.v_bias = nn.Parameter(torch.zeros(d))


class VideoMAESelfOutput(_Holder):
    def __init__(self, d):
        super().__init__()
        self.dense = nn.Linear(d, d)


class VideoMAEAttention(_Holder):
    def __init__(self, d):
        super().__init__()
        self.attention = VideoMAESelfAttention(d)
        self.output = VideoMAESelfOutput(d)


class _Dense(_Holder):
    def __init__(self, i, o):
        super().__init__()
        self.dense = nn.Linear(i, o)


class VideoMAELayer(_Holder):
    def __init__(self, d, ff, eps):
        super().__init__()
        self.attention = VideoMAEAttention(d)
        self.intermediate = _Dense(d, ff)
        self.output = _Dense(ff, d)
        self.layernorm_before = nn.LayerNorm(d, eps=eps)
        self.layernorm_after = nn.LayerNorm(d, eps=eps)

    def flat_params(self):
        a = self.attention.attention
        return (self.layernorm_before.weight, self.layernorm_before.bias, a.query.weight, a.key.weight, a.value.weight,
                a.q_bias, a.v_bias, self.attention.output.dense.weight, self.attention.output.dense.bias,
                self.layernorm_after.weight, self.layernorm_after.bias, self.intermediate.dense.weight,
                self.intermediate.dense.bias, self.output.dense.weight, self.output.dense.bias)


class VideoMAEEncoder(_Holder):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList(VideoMAELayer(c.hidden_size, c.intermediate_size, c.layer_norm_eps)
                                   for _ in range(c.num_hidden_layers))


class VideoMAEModel(_Holder):
    def __init__(self, c):
        super().__init__()
        self.embeddings = VideoMAEEmbeddings(c)
        self.encoder = VideoMAEEncoder(c)
        # use_mean_pooling=True -> no final LayerNorm (HF:415-418); the reference always sets it (pretrain_videomae.py:55)


class VideoMAEDecoder(_Holder):
    def __init__(self, c):
        super().__init__()
        self.decoder_layers = nn.ModuleList(
            VideoMAELayer(c.decoder_hidden_size, c.decoder_intermediate_size, c.layer_norm_eps)
            for _ in range(c.decoder_num_hidden_layers))
        self.norm = nn.LayerNorm(c.decoder_hidden_size)  # eps 1e-5 (HF:497)
        self.head = nn.Linear(c.decoder_hidden_size, c.num_channels * c.tubelet_size * c.patch_size ** 2)


class VideoMAEForPreTraining(nn.Module):
    """`model(pixel_values, bool_masked_pos=...) -> VideoMAEForPreTrainingOutput(loss, logits)`  (HF:540-680)."""

    def __init__(self, config, output_logits: bool = True):
        super().__init__()
        c = config
        if not getattr(c, "use_mean_pooling", True):
            raise NotImplementedError("use_mean_pooling=False (final encoder LayerNorm) is not on the reference's path")
        if getattr(c, "hidden_act", "gelu") != "gelu" or not getattr(c, "qkv_bias", True):
            raise NotImplementedError("only hidden_act='gelu' and qkv_bias=True (the reference's configuration)")
        if c.hidden_size % c.num_attention_heads or c.hidden_size // c.num_attention_heads != 64 or \
                c.decoder_hidden_size // c.decoder_num_attention_heads != 64:
            raise NotImplementedError("the sm_100a attention kernels are built for head_dim 64")
        if c.num_channels != 3 or c.patch_size != 16 or c.tubelet_size not in (1, 2):
            raise NotImplementedError("patchify kernel: 3 channels, 16x16 patches, tubelet 1 or 2")
        if max(c.hidden_size, c.decoder_hidden_size) > 1024 or c.image_size > 256 or c.image_size % c.patch_size or \
                c.num_frames % c.tubelet_size:
            raise NotImplementedError("LayerNorm kernels keep a row in registers (width <= 1024: up to ViT-L); the "
                                      "patchify kernel takes frames of whole 16x16 patches up to 256 pixels wide and "
                                      "whole tubelets")
        self.config = c
        self.output_logits = output_logits
        self.videomae = VideoMAEModel(c)
        self.encoder_to_decoder = nn.Linear(c.hidden_size, c.decoder_hidden_size, bias=False)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, c.decoder_hidden_size))
        self.decoder = VideoMAEDecoder(c)
        g = c.image_size // c.patch_size
        self.num_patches = (c.num_frames // c.tubelet_size) * g * g
        # fixed sin-cos tables: plain attributes, NOT buffers (absent from the state-dict, as in HF:104, HF:529)
        self.position_embeddings_encoder = get_sinusoid_encoding_table(self.num_patches, c.hidden_size)
        self.position_embeddings = get_sinusoid_encoding_table(self.num_patches, c.decoder_hidden_size)
        self._pos_dev = {}
        self._cache = Bf16Cache()
        self._grad_sync = None   # set by bvc_b200.DistributedDataParallel (ddp.py)
        self._stage_param_map = None
        self._pixel_norm = None  # (mean[3], std[3]) for uint8 pixel_values, see set_input_normalization
        self._nv = None          # visible-token count of the current call
        self._nv_cache = {}      # static_mask_count: (B, N, device) -> count read back once
        self._status = None      # device int32 flag: a mask row violated the equal-count contract
        self._status_host = self._status_event = None
        self.mask_mismatches = 0
        # True: trust the first call's visible-token count for later calls of the same shape (no per-call readback of
        # a CUDA mask; validated on the device).  Default: every call counts its own mask, like HF.
        self.static_mask_count = os.environ.get("BVC_STATIC_MASK_COUNT", "0") == "1"
        self._init_weights(c.initializer_range)

    # -- init: modeling_utils.py:2285-2325 (normal(0, initializer_range) weights, zero biases, LN 1/0) -------------
    def _init_weights(self, std):
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Conv3d)):
                nn.init.normal_(m.weight, mean=0.0, std=std)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.LayerNorm):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, VideoMAESelfAttention):
                nn.init.zeros_(m.q_bias)
                nn.init.zeros_(m.v_bias)
        nn.init.zeros_(self.mask_token)

    def _pos(self, dev):
        t = self._pos_dev.get(dev)
        if t is None:
            t = (self.position_embeddings_encoder.to(dev).contiguous(), self.position_embeddings.to(dev).contiguous())
            self._pos_dev[dev] = t
        return t

    def _weight_entries(self):
        ents = [("pe", "w", (self.videomae.embeddings.patch_embeddings.projection.weight,))]
        for tag, layers in (("e", self.videomae.encoder.layer), ("d", self.decoder.decoder_layers)):
            for i, layer in enumerate(layers):
                a = layer.attention.attention
                n = f"{tag}{i}."
                ents.append((n + "qkv", "qkv", (a.query.weight, a.key.weight, a.value.weight, a.q_bias, a.v_bias)))
                ents.append((n + "wo", "w", (layer.attention.output.dense.weight,)))
                ents.append((n + "w1", "w", (layer.intermediate.dense.weight,)))
                ents.append((n + "w2", "w", (layer.output.dense.weight,)))
        ents.append(("e2d", "w", (self.encoder_to_decoder.weight,)))
        ents.append(("head", "w", (self.decoder.head.weight,)))
        return ents

    def _stage_params(self):
        """stage name -> parameters in the order that stage's backward returns their gradients (engine.py); the
        gradient synchroniser checks each .grad against the reduced view it was handed (ddp.GradSync)."""
        if self._stage_param_map is None:
            proj = self.videomae.embeddings.patch_embeddings.projection
            m = {"pe": (proj.weight, proj.bias), "e2d": (self.encoder_to_decoder.weight, self.mask_token),
                 "head": (self.decoder.norm.weight, self.decoder.norm.bias, self.decoder.head.weight,
                          self.decoder.head.bias)}
            for tag, layers in (("e", self.videomae.encoder.layer), ("d", self.decoder.decoder_layers)):
                for i, layer in enumerate(layers):
                    m[f"{tag}{i}."] = layer.flat_params()
            self._stage_param_map = m
        return self._stage_param_map

    def set_input_normalization(self, mean, std):
        """Accept uint8 clips: `model(pixel_values_uint8, ...)` then applies the dataset's ToTensor + Normalize
        (pretraining/generative/homeview.py:218-231, `Normalize(mean, std)` after `/ 255`) inside the patchify kernel --
        bit-identical to normalising on the host, at a quarter of the host->device and HBM bytes.
        The reference's transform is mean = std-per-channel constants (0.5, 0.25)."""
        mean, std = tuple(float(v) for v in mean), tuple(float(v) for v in std)
        if len(mean) != 3 or len(std) != 3 or any(v == 0.0 for v in std):
            raise ValueError("mean / std must have 3 entries, std non-zero")
        self._pixel_norm = (mean, std)
        return self

    def weight_shadows(self):
        """bf16 / packed-bias operand copies of the weights, for an optimizer that refreshes them in its own pass."""
        return self._cache.shadows()

    def weight_shadows_synced(self):
        self._cache.mark_synced()

    def _visible_count(self, mask_in, m_dev, B, N, dev):
        """Visible tokens per row (HF:121-122 reshapes `embeddings[~mask]` to [B, -1, C]: every row must mask the same
        number of tokens, and any such number is legal on any call).  A mask that still lives on the HOST -- where the
        reference builds it, pretrain_videomae.py:293-297 -- is counted there: exact, no device round trip.  A CUDA mask
        is counted on the device and read back (one small synchronising copy per call; the reference's own
        `bool_masked_pos.to(rank)` from pageable memory synchronises the stream as well).  `static_mask_count = True`
        opts out: the count is read back once per (batch, tokens) shape, later calls are validated on the device only
        (mismatch -> NaN loss for that call, detected and recovered from on the next call, surfaced by
        check_mask_status())."""
        if not mask_in.is_cuda:
            mh = mask_in if mask_in.dtype == torch.bool else mask_in != 0
            cnt = (~mh).sum(dim=1)
            if not bool((cnt == cnt[0]).all()):
                raise ValueError("bool_masked_pos: every row must mask the same number of tokens (HF:121-122 reshape)")
            return int(cnt[0])
        key = (B, N, str(dev))
        if self.static_mask_count:
            # (under CUDA-graph capture no event may be queried: the status copy of a replayed step is checked by the
            # next eager call, or by check_mask_status())
            capturing = torch.cuda.is_current_stream_capturing()
            if not capturing and self._status_event is not None and self._status_event.query() \
                    and int(self._status_host[0]) != 0:
                # an earlier call's rows did not match the cached count: forget it and count this mask for real
                self._status.zero_()
                self._status_host.zero_()
                self._nv_cache.clear()
                self.mask_mismatches += 1
            nv = self._nv_cache.get(key)
            if nv is not None:
                return nv
        cnt = torch.empty(B, dtype=torch.int32, device=dev)
        L.mask_count(m_dev, cnt)
        cnt = cnt.cpu()
        if not bool((cnt == cnt[0]).all()):
            raise ValueError("bool_masked_pos: every row must mask the same number of tokens (HF:121-122 reshape)")
        nv = int(cnt[0])
        if self.static_mask_count:
            self._nv_cache[key] = nv
        return nv

    def check_mask_status(self):
        """Synchronise and raise if any forward since the last check saw rows with unequal mask counts."""
        if self._status is not None and int(self._status.item()) != 0:
            self._status.zero_()
            raise ValueError("bool_masked_pos: every row must mask the same number of tokens (HF:121-122 reshape)")

    # -- the un-masked encoder pass ---------------------------------------------------------------------------------
    def encode(self, pixel_values):
        """`VideoMAEModel.forward(pixel_values, bool_masked_pos=None).last_hidden_state` (HF:420-470 with
        use_mean_pooling=True, i.e. no final LayerNorm): ALL N tokens through the patch embedding (+ sinusoid table) and
        the encoder blocks -- the pass benchmarks/compute_embeddings_videomae.py:261 runs on a pretrained checkpoint
        (its VideoMAEForVideoClassification head -- mean over tokens, fc_norm, classifier -- stays the caller's).
        Same kernels as the training step at S = N (1568) tokens, 12 heads.  -> fp32 [B, N, hidden]."""
        c = self.config
        if pixel_values.dim() != 5:
            raise ValueError("pixel_values must be [batch, frames, channels, height, width]")
        B, T, C, H, W = pixel_values.shape
        if C != c.num_channels:
            raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the "
                             "configuration.")
        if H != c.image_size or W != c.image_size or T != c.num_frames:
            raise ValueError(f"expected [{c.num_frames}, {c.num_channels}, {c.image_size}, {c.image_size}] clips")
        if not pixel_values.is_cuda:
            raise L.BvcError("VideoMAEForPreTraining (bvc-b200) runs on CUDA only; there is no CPU path")
        dev = pixel_values.device
        N = self.num_patches
        x = pixel_values.detach()
        if x.dtype == torch.uint8:
            if self._pixel_norm is None:
                raise ValueError("uint8 pixel_values need model.set_input_normalization(mean, std) first")
            x = x.contiguous()
        elif x.dtype != F32 or not x.is_contiguous():
            x = x.to(F32).contiguous()
        with torch.cuda.device(dev):
            st = StepState(self._cache)
            st.B, st.N, st.nv, st.nm = B, N, N, 0
            st.status = torch.zeros(1, dtype=torch.int32, device=dev)
            st.vis_idx = torch.empty((B, N), dtype=torch.int32, device=dev)
            st.msk_idx = torch.empty((1,), dtype=torch.int32, device=dev)
            st.slot = torch.empty((B, N), dtype=torch.int32, device=dev)
            L.mask_to_index(torch.zeros((B, N), dtype=torch.uint8, device=dev), N, st.vis_idx, st.msk_idx, st.slot,
                            st.status)
            K = C * c.tubelet_size * c.patch_size ** 2
            patches = torch.empty((B * N, K), dtype=BF16, device=dev)
            L.patchify_target(x, st.slot, c.tubelet_size, c.patch_size, N, patches,
                              torch.empty((1, K), dtype=F32, device=dev), False,
                              pixel_norm=self._pixel_norm if x.dtype == torch.uint8 else None)
            pos_e, _ = self._pos(dev)
            proj = self.videomae.embeddings.patch_embeddings.projection
            self._cache.register(self._weight_entries(), dev)
            self._cache.refresh()
            st.epoch = self._cache.epoch
            h = EmbedFn.apply(proj.weight, proj.bias, patches, pos_e, st, "pe")
            for i, layer in enumerate(self.videomae.encoder.layer):
                h = BlockFn.apply(h, *layer.flat_params(), st, f"e{i}.", B, N, c.num_attention_heads,
                                  float(c.layer_norm_eps))
        return h.view(B, N, c.hidden_size)

    # -- the hot path -------------------------------------------------------------------------------------------------
    def forward(self, pixel_values, bool_masked_pos=None, **kwargs):
        c = self.config
        if pixel_values.dim() != 5:
            raise ValueError("pixel_values must be [batch, frames, channels, height, width]")
        B, T, C, H, W = pixel_values.shape
        if C != c.num_channels:
            raise ValueError("Make sure that the channel dimension of the pixel values match with the one set in the "
                             "configuration.")  # HF:166-169
        if H != c.image_size or W != c.image_size:
            raise ValueError(f"Input image size ({H}*{W}) doesn't match model ({c.image_size}*{c.image_size}).")
        if T != c.num_frames:
            raise ValueError(f"expected {c.num_frames} frames, got {T}")
        if bool_masked_pos is None:
            raise ValueError("One must provided a boolean mask ")  # HF:582-583
        if not pixel_values.is_cuda:
            raise L.BvcError("VideoMAEForPreTraining (bvc-b200) runs on CUDA only; there is no CPU path")
        dev = pixel_values.device
        N = self.num_patches
        if tuple(bool_masked_pos.shape) != (B, N):
            raise ValueError(f"bool_masked_pos must be [{B}, {N}]")
        x = pixel_values.detach()
        if x.dtype == torch.uint8:
            if self._pixel_norm is None:
                raise ValueError("uint8 pixel_values need model.set_input_normalization(mean, std) first")
            x = x.contiguous()
        elif x.dtype != F32 or not x.is_contiguous():
            x = x.to(F32).contiguous()
        m = bool_masked_pos.to(device=dev, non_blocking=True)
        m = (m if m.dtype == torch.bool else m != 0).contiguous().view(torch.uint8)

        with torch.cuda.device(dev):
            if self._status is None or self._status.device != dev:
                self._status = torch.zeros(1, dtype=torch.int32, device=dev)
                self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory()
                self._status_event = None
            self._nv = self._visible_count(bool_masked_pos, m, B, N, dev)
            nv = self._nv
            nm = N - nv
            if nv == 0 or nm == 0:
                raise ValueError("bool_masked_pos must leave at least one visible and one masked token")

            st = StepState(self._cache)
            st.B, st.N, st.nv, st.nm, st.status = B, N, nv, nm, self._status
            st.sync = self._grad_sync
            if st.sync is not None:
                st.stage_params = self._stage_params()
            st.vis_idx = torch.zeros((B, nv), dtype=torch.int32, device=dev)
            st.msk_idx = torch.zeros((B, nm), dtype=torch.int32, device=dev)
            st.slot = torch.empty((B, N), dtype=torch.int32, device=dev)
            L.mask_to_index(m, nv, st.vis_idx, st.msk_idx, st.slot, st.status)
            if self.static_mask_count:
                self._status_host.copy_(self._status, non_blocking=True)
                if not torch.cuda.is_current_stream_capturing():
                    self._status_event = torch.cuda.Event()
                    self._status_event.record()

            K = C * c.tubelet_size * c.patch_size ** 2
            patches = torch.empty((B * nv, K), dtype=BF16, device=dev)
            target = torch.empty((B * nm, K), dtype=F32, device=dev)
            L.patchify_target(x, st.slot, c.tubelet_size, c.patch_size, nv, patches, target,
                              bool(getattr(c, "norm_pix_loss", True)),
                              pixel_norm=self._pixel_norm if x.dtype == torch.uint8 else None)

            pos_e, pos_d = self._pos(dev)
            proj = self.videomae.embeddings.patch_embeddings.projection
            self._cache.register(self._weight_entries(), dev)
            self._cache.refresh()
            st.epoch = self._cache.epoch
            h = EmbedFn.apply(proj.weight, proj.bias, patches, pos_e, st, "pe")
            for i, layer in enumerate(self.videomae.encoder.layer):
                h = BlockFn.apply(h, *layer.flat_params(), st, f"e{i}.", B, nv, c.num_attention_heads,
                                  float(c.layer_norm_eps))
            xf = EncToDecFn.apply(h, self.encoder_to_decoder.weight, self.mask_token, pos_d, st, "e2d")
            for j, layer in enumerate(self.decoder.decoder_layers):
                xf = BlockFn.apply(xf, *layer.flat_params(), st, f"d{j}.", B, N, c.decoder_num_attention_heads,
                                   float(c.layer_norm_eps))
            loss = HeadLossFn.apply(xf, self.decoder.norm.weight, self.decoder.norm.bias, self.decoder.head.weight,
                                    self.decoder.head.bias, target, st, "head")
            logits_fn = None
            if self.output_logits:
                z, whb, bh, bn = st.head_operands
                st.head_operands = None

                def logits_fn(z=z, whb=whb, bh=bh.detach(), bn=bn, st=st, shape=(B, nm, K), dev=dev):
                    st.check()  # the bf16 weight copies must still be the ones this forward used
                    with torch.cuda.device(dev):
                        out = torch.empty((shape[0] * shape[1], shape[2]), dtype=BF16, device=dev)
                        L.gemm(z, whb, out.shape[0], shape[2], z.shape[1], out_bf16=out, bias=bh, block_n=bn)
                    return out.view(shape)
        return VideoMAEForPreTrainingOutput(loss=loss, logits_fn=logits_fn)
