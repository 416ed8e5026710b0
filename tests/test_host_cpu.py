"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/bvc.h declares, the host-side
mirrors of the reference interface behave like the reference, and the product path refuses to run without CUDA."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import videomae_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bvc():
    import __graft_entry__ as g
    g.build()
    import bvc_b200
    return bvc_b200


def test_library_exports_every_declared_symbol(bvc):
    hdr = open(os.path.join(ROOT, "include", "bvc.h")).read()
    declared = set(re.findall(r"\b(bvc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = bvc._lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(bvc._lib.EXPORTED_SYMBOLS)
    assert lib.bvc_abi_version() == bvc._lib.ABI_VERSION
    # host-only entry point: tile bookkeeping of the fused loss epilogue
    assert lib.bvc_gemm_loss_slots(256, 512, 256) == 2 * 2 * 8
    assert lib.bvc_smooth_l1_slots(1) == 1 and lib.bvc_smooth_l1_slots(10 ** 9) > 1


def test_plain_c_host_compiles_and_links_against_the_header(bvc, tmp_path):
    """"a C/C++ host links the same way".  include/bvc.h must be valid C (not only C++), and a host
    program that references EVERY declared entry point must link against libbvc.so and run without a GPU (nothing is
    launched: only bvc_abi_version and the host-only bookkeeping calls are executed)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    hdr = open(os.path.join(ROOT, "include", "bvc.h")).read()
    declared = sorted(set(re.findall(r"\b(bvc_[a-z0-9_]+)\s*\(", hdr)))
    table = ",\n".join(f"  (void (*)(void)){n}" for n in declared)
    src = tmp_path / "host.c"
    src.write_text(
        '#include <stdio.h>\n#include "bvc.h"\n'
        f"static void (*const entry[])(void) = {{\n{table}\n}};\n"
        "int main(void) {\n"
        "  unsigned n = 0, i;\n"
        "  for (i = 0; i < sizeof entry / sizeof entry[0]; ++i) n += entry[i] != 0;\n"
        '  printf("%d %u %ld\\n", bvc_abi_version(), n, (long)bvc_gemm_loss_slots(256, 512, 256));\n'
        "  return bvc_abi_version() == BVC_ABI_VERSION ? 0 : 1;\n}\n")
    libdir = os.path.dirname(bvc._lib.LIB_PATH)
    exe = tmp_path / "host"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lbvc", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert out == [str(bvc._lib.ABI_VERSION), str(len(declared)), "32"]


def test_masking_mirror_matches_reference_fixtures(bvc, golden_dir):
    g = np.load(os.path.join(golden_dir, "masks.npz"))
    np.random.seed(0)
    gen = bvc.TubeMaskingGenerator((8, 14, 14), 0.9)
    got = np.stack([gen() for _ in range(4)]).astype(np.uint8)
    assert np.array_equal(got, g["tube_8x14x14_r90"])
    assert gen.total_masks == 1408 and gen.num_masks_per_frame == 176 and "mask patches 1408" in repr(gen)
    np.random.seed(0)
    rg = bvc.RandomMaskingGenerator((8, 14, 14), 0.9)
    assert np.array_equal(np.stack([rg() for _ in range(2)]).astype(np.uint8), g["random_8x14x14_r90"])
    np.random.seed(5)
    m = bvc.batch_masks(bvc.TubeMaskingGenerator((8, 14, 14), 0.9), 3)
    assert m.dtype == torch.bool and m.shape == (3, 1568) and bool((m.sum(1) == 1408).all())


def test_module_matches_hf_state_dict_contract(bvc):
    for name in ("tiny", "small", "base"):
        cfg = O.make_config(name)
        from tests.helpers import bvc_config
        m = bvc.VideoMAEForPreTraining(bvc_config(cfg))
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert shapes == {k: tuple(s) for k, s in O.param_shapes(cfg).items()}
        assert not list(m.buffers())  # sinusoid tables are plain attributes, as in HF
        assert m.config.image_size == cfg.image_size and m.config.tubelet_size == cfg.tubelet_size
        # HF init: zero biases / mask_token / q_bias / v_bias, unit LayerNorm, N(0, 0.02) weights
        sd = m.state_dict()
        assert float(sd["mask_token"].abs().max()) == 0.0
        assert float(sd["decoder.head.bias"].abs().max()) == 0.0
        assert abs(float(sd["decoder.head.weight"].std()) - 0.02) < 2e-3
        assert torch.equal(m.position_embeddings, O.sinusoid_table(cfg.seq_len, cfg.decoder_hidden_size))


def test_hf_state_dict_loads_both_ways(bvc):
    transformers = pytest.importorskip("transformers")
    c = transformers.VideoMAEConfig(image_size=32, num_frames=4, hidden_size=64, num_hidden_layers=2,
                                    num_attention_heads=1, intermediate_size=128, decoder_num_attention_heads=1,
                                    decoder_hidden_size=64, decoder_num_hidden_layers=1,
                                    decoder_intermediate_size=128, norm_pix_loss=True)
    hf = transformers.VideoMAEForPreTraining(c)
    ours = bvc.VideoMAEForPreTraining(c)  # a real transformers config object is accepted
    ours.load_state_dict(hf.state_dict(), strict=True)
    hf.load_state_dict(ours.state_dict(), strict=True)
    assert [k for k, _ in ours.named_parameters()] == [k for k, _ in hf.named_parameters()]


def test_no_cpu_fallback(bvc):
    cfg = O.make_config("tiny")
    from tests.helpers import bvc_config
    m = bvc.VideoMAEForPreTraining(bvc_config(cfg))
    x = O.synthetic_clip(1, cfg)
    mask = torch.zeros(1, cfg.seq_len, dtype=torch.bool)
    mask[:, :4] = True
    with pytest.raises(ValueError):
        m(x)  # mask is mandatory (HF:582-583) -- checked before any device work
    with pytest.raises(bvc.BvcError):
        m(x, bool_masked_pos=mask)
    with pytest.raises(bvc.BvcError):
        bvc._lib.cast_bf16(torch.zeros(8), torch.zeros(8, dtype=torch.bfloat16))


def test_unsupported_configs_fail_loudly(bvc):
    with pytest.raises(NotImplementedError):
        bvc.VideoMAEForPreTraining(bvc.VideoMAEConfig(num_attention_heads=8))  # head_dim 96
    with pytest.raises(NotImplementedError):
        bvc.VideoMAEForPreTraining(bvc.VideoMAEConfig(use_mean_pooling=False))
    # limits of the kernels (include/bvc.h) surface at construction, not as a failed launch in the first forward
    for kw in (dict(hidden_size=1280, num_attention_heads=20, intermediate_size=5120), dict(image_size=384),
               dict(image_size=200), dict(num_frames=15), dict(patch_size=14), dict(tubelet_size=4)):
        with pytest.raises(NotImplementedError):
            bvc.VideoMAEForPreTraining(bvc.VideoMAEConfig(**kw))


def test_bench_flop_model_matches_baseline_md():
    import bench
    assert abs(bench.flops_per_clip(bench.CONFIGS["base"]) / 1e9 - 202.3) < 0.5
    assert abs(bench.flops_per_clip(bench.CONFIGS["small"]) / 1e9 - 64.0) < 0.5
    assert abs(bench.flops_per_clip(bench.CONFIGS["large"]) / 1e9 - 484.4) < 1.0


def test_fused_sgd_host_side():
    """FusedSGD mirrors torch.optim.SGD's constructor checks and refuses CPU parameters (no CPU fallback)."""
    import pytest
    import torch
    import bvc_b200 as bvc
    p = torch.nn.Parameter(torch.zeros(8))
    with pytest.raises(ValueError):
        bvc.FusedSGD([p], lr=-1.0)
    with pytest.raises(ValueError):
        bvc.FusedSGD([p], lr=0.1, nesterov=True)  # nesterov needs momentum (torch/optim/sgd.py)
    opt = bvc.FusedSGD([p], lr=0.1, momentum=0.9, nesterov=True)
    assert opt._step_supports_amp_scaling and opt.defaults["nesterov"]
    p.grad = torch.ones(8)
    with pytest.raises(bvc.BvcError):
        opt.step()


def test_simclr_host_side(golden_dir):
    """get_special_matrix / make_masks mirror pretrain_simclr.py:86-91, 285-291 bit-exactly (fixture from the reference's
    own function); info_nce_loss refuses CPU tensors (no CPU fallback) and bad shapes."""
    import os
    import numpy as np
    import pytest
    import torch
    import bvc_b200 as bvc
    g = np.load(os.path.join(golden_dir, "simclr_tiny.npz"))
    assert np.array_equal(bvc.get_special_matrix(16), g["special"])
    pos, neg = bvc.make_simclr_masks(16, "cpu")
    assert pos.dtype == torch.bool and int(pos.sum()) == 30 and int(neg.sum()) == 16 * 16 - 16 - 30
    assert not bool((pos & neg).any()) and not bool(neg.diagonal().any())
    with pytest.raises(bvc.BvcError):
        bvc.info_nce_loss(0.1, (pos, neg), torch.zeros(16, 32))
    with pytest.raises(ValueError):
        bvc.info_nce_loss(0.1, (pos, neg), torch.zeros(16))


def test_jepa_host_mirror_validates_before_touching_cuda():
    """bvc_b200.apply_masks / repeat_interleave_batch / smooth_l1_loss / ema_update (jepa.py): argument errors are Python
    ValueErrors like the reference's torch calls would raise; CPU tensors never reach a kernel (BvcError, no fallback)."""
    import pytest
    import torch
    import bvc_b200 as bvc
    x = torch.randn(2, 10, 8)
    with pytest.raises(ValueError):
        bvc.apply_masks(x, [])
    with pytest.raises(ValueError):
        bvc.apply_masks(x, [torch.zeros(2, 3, dtype=torch.int64), torch.zeros(2, 4, dtype=torch.int64)])
    with pytest.raises(ValueError):
        bvc.apply_masks(torch.randn(10, 8), [torch.zeros(2, 3, dtype=torch.int64)])
    with pytest.raises(ValueError):
        bvc.apply_masks(x.half(), [torch.zeros(2, 3, dtype=torch.int64)])   # fp16 is not a dtype of this path
    with pytest.raises(ValueError):
        bvc.repeat_interleave_batch(torch.randn(5, 8), 2, 2)
    with pytest.raises(ValueError):
        bvc.smooth_l1_loss(torch.randn(4, 4), torch.randn(4, 5))
    with pytest.raises(ValueError):
        bvc.ema_update([torch.randn(3)], [], 0.99)
    with pytest.raises(ValueError):
        bvc.jepa_targets(torch.randn(2, 10, 6), [torch.zeros(2, 3, dtype=torch.int64)], 1)   # D % 4 != 0
    with pytest.raises(bvc.BvcError):
        bvc.apply_masks(x, [torch.zeros(2, 3, dtype=torch.int64)])


def test_graphed_train_step_host_side(bvc):
    """GraphedTrainStep validates its arguments without touching CUDA, switches the model to the device-validated
    visible-token count, and (like every other entry point) has no CPU path."""
    from tests.helpers import bvc_config
    cfg = O.make_config("tiny")
    m = bvc.VideoMAEForPreTraining(bvc_config(cfg))
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    assert m.static_mask_count is False
    step = bvc.GraphedTrainStep(m, opt, None, warmup=2)
    assert m.static_mask_count is True and (step.calls, step.captures, step.replays) == (0, 0, 0)
    with pytest.raises(ValueError):
        bvc.GraphedTrainStep(m, opt, None, warmup=1)
    with pytest.raises(TypeError):
        bvc.GraphedTrainStep(torch.nn.Linear(2, 2), opt, None)
    x = O.synthetic_clip(1, cfg, seed=0)
    np.random.seed(0)
    mask = O.batch_tube_masks(1, cfg.grid, 0.5)
    with pytest.raises(bvc.BvcError):  # CPU tensors: the eager warm-up call reaches the model, which has no CPU path
        step(x, mask)


def test_gelu_fit_in_the_gemm_epilogue_is_exact_erf_gelu_to_fp32_noise():
    """HF:316 is the exact-erf GELU.  The GEMM epilogue evaluates Phi(-|x|) = 2^Q(|x|) with a degree-7 polynomial
    (csrc/gemm_kernel.cuh, fitted by tools/fit_gelu.py): read the coefficients out of the kernel source, evaluate them in
    fp32 exactly as the kernel does (Horner with FMAs, |x| clamped at 6) and compare gelu / gelu' with scipy's erf over
    the whole bf16 input range.  Bounds: relative error of Phi(-a) <= 1e-5 on [0, 6] (3.3e-6 of fit error plus the
    fp32 rounding of Q near -30 in the tail); gelu and gelu' within 4e-6
    absolute of the exact values everywhere -- three orders of magnitude below the bf16 rounding of the outputs
    (2^-9 relative), so the fit cannot show in any parity figure."""
    from scipy import special
    src = open(os.path.join(ROOT, "baby-vision-curriculum_b200", "csrc", "gemm_kernel.cuh")).read()
    body = src[src.index("float phi_neg_abs(float x)"):src.index("return exp2f(q);")]
    coef = [float(c) for c in re.findall(r"(-?\d\.\d+e[+-]\d+)f", body)]
    assert len(coef) == 8, coef
    # the packed fp32x2 variant of the fast path must carry the same coefficients
    body2 = src[src.index("uint64_t phi_neg_abs2(uint64_t a2)"):src.index("unpack2(q, q0, q1);")]
    coef2 = [float(c) for c in re.findall(r"pack2\((-?\d\.\d+e[+-]\d+)f,", body2)]
    assert coef2 == coef

    def phi_neg_abs(x):
        a = np.minimum(np.abs(x), np.float32(6.0)).astype(np.float32)
        q = np.full_like(a, np.float32(coef[0]))
        for c in coef[1:]:
            q = (q.astype(np.float64) * a + np.float32(c)).astype(np.float32)  # one rounding per step, like fmaf
        return np.exp2(q.astype(np.float64)).astype(np.float32)

    a = np.linspace(0, 6, 600001).astype(np.float32)
    rel = np.abs(phi_neg_abs(a).astype(np.float64) / special.ndtr(-a.astype(np.float64)) - 1)
    assert rel.max() <= 1e-5, rel.max()
    # every finite bf16 value in [-60, 60] (beyond |x| = 6 the clamp leaves Phi(-|x|) = 1e-9: gelu is x or 0 to 1e-7)
    bits = np.arange(0, 1 << 16, dtype=np.uint32) << 16
    x = bits.view(np.float32)
    x = x[np.isfinite(x) & (np.abs(x) <= 60)]
    w = phi_neg_abs(x)
    xw = x * w
    gelu = np.where(x > 0, x - xw, xw).astype(np.float64)
    x64 = x.astype(np.float64)
    assert np.abs(gelu - x64 * special.ndtr(x64)).max() <= 4e-6
    cdf = np.where(x > 0, 1.0 - w, w).astype(np.float64)
    pdf = 0.3989422804014327 * np.exp2(-0.72134752044448170 * x64 * x64)
    grad = x64 * pdf + cdf
    exact = special.ndtr(x64) + x64 * np.exp(-0.5 * x64 * x64) / np.sqrt(2 * np.pi)
    assert np.abs(grad - exact).max() <= 4e-6


def test_every_entry_point_rejects_null_arguments_before_touching_cuda(bvc):
    """Error behaviour of the C ABI (include/bvc.h: "returns 0 on success, <0 on error; never throws"): every entry
    point called with null pointers and zero sizes returns BVC_ERR_ARG from its host-side validation -- no crash, no
    CUDA call (this runs without a GPU), one diagnostic line on stderr.  In a child process, so that a crash would be a
    test failure and not the end of the test session."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from bvc_b200 import _lib as L
lib = L.load()
n = 0
for name, (res, args) in L._SIGNATURES.items():
    if res is not C.c_int or not args:
        continue
    vals = [0 if a in (C.c_int32, C.c_int64, C.c_int) else 0.0 if a in (C.c_float, C.c_double) else None for a in args]
    rc = getattr(lib, name)(*vals)
    assert rc == -1, (name, rc)
    n += 1
print("REJECTED", n)
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    n_int = sum(1 for res, args in bvc._lib._SIGNATURES.values() if res is __import__("ctypes").c_int and args)
    assert r.stdout.split() == ["REJECTED", str(n_int)] and n_int >= 29
    assert r.stderr.count("bvc: bad argument") == n_int


def test_gemm_entry_point_argument_contract(bvc):
    """bvc_gemm_bf16's documented preconditions (include/bvc.h), each violated alone with otherwise valid arguments and
    fake, never dereferenced device addresses: BVC_ERR_ARG before any CUDA call."""
    import ctypes as C
    lib = bvc._lib.load()
    P = 0x10000  # 16-byte aligned fake device address

    def call(**over):
        kw = dict(a=P, b=P + 0x1000, lda=64, ldb=64, a_mn_major=0, b_mn_major=0, M=128, N=64, K=64, k_splits=1,
                  out_f32=None, out_bf16=P + 0x2000, ldo=64, alpha_host=1.0, act=0, ld_aux=0, ldr=0, ldt=0,
                  block_n=0, cta_pair=0)
        kw.update(over)
        args = bvc._lib.GemmArgs(**kw)
        return lib.bvc_gemm_bf16(C.byref(args), None)

    bad = [dict(a=None), dict(b=None), dict(M=0), dict(K=-1), dict(N=60), dict(lda=60), dict(ldb=68), dict(ldo=4),
           dict(block_n=100), dict(cta_pair=3), dict(out_bf16=None), dict(a=P + 8), dict(b=P + 0x1004),
           dict(act=3), dict(act=2), dict(act=1, ld_aux=12), dict(res=P, ldr=6),
           dict(target=P, ldt=64, loss_partial=None), dict(target=P, ldt=6, loss_partial=P),
           dict(colsum=P, k_splits=2), dict(colsum=P, out_seg=16), dict(colsum=P, target=P, ldt=64, loss_partial=P),
           dict(lda=32), dict(ldb=32), dict(a_mn_major=1, lda=64, M=128), dict(b_mn_major=1, ldb=32, N=64),
           dict(cta_pair=2, block_n=192, b_mn_major=1, ldb=256, N=256), dict(cta_pair=2, block_n=64)]
    for over in bad:
        assert call(**over) == -1, over


def test_patchify_layernorm_attention_argument_contracts(bvc):
    """The same for the other entry points with documented limits (include/bvc.h): 3 channels, 16 x 16 patches, tubelet
    1 or 2, W <= 256 and whole patches for the patchify pass; d % 4 == 0, d <= the register-resident row limit and
    ld >= d for LayerNorm; 16-byte aligned operands for attention; a zero std for the uint8 path."""
    import ctypes as C
    lib = bvc._lib.load()
    P = 0x10000

    def patchify(pixels=P, B=2, T=16, Cc=3, H=224, W=224, ts=2, ps=16, nv=160):
        return lib.bvc_patchify_target(pixels, P, B, T, Cc, H, W, ts, ps, nv, P, P, 1, None)

    for kw in (dict(Cc=4), dict(ps=8), dict(ts=3), dict(ts=4), dict(T=15), dict(H=220), dict(W=232), dict(W=272),
               dict(B=0), dict(pixels=P + 4), dict(nv=1569), dict(nv=-1), dict(pixels=None)):
        assert patchify(**kw) == -1, kw
    f3 = C.c_float * 3
    u8 = lambda mean, std: lib.bvc_patchify_target_u8(P, mean, std, P, 2, 16, 3, 224, 224, 2, 16, 160, P, P, 1, None)  # noqa: E731
    assert u8(f3(0.5, 0.5, 0.5), f3(0.25, 0.0, 0.25)) == -1      # std 0: the division of homeview.py:222 is undefined
    assert u8(None, f3(0.25, 0.25, 0.25)) == -1 and u8(f3(0.5, 0.5, 0.5), None) == -1

    def ln(d=768, ldx=768, M=64):
        return lib.bvc_layernorm_fwd(P, ldx, 0, 0, 0, P, P, 1e-12, M, d, P, P, P, None)

    for kw in (dict(d=770), dict(d=0), dict(M=0), dict(ldx=512), dict(ldx=770), dict(d=1 << 20, ldx=1 << 20)):
        assert ln(**kw) == -1, kw
    assert lib.bvc_attn_fwd(P + 8, 2, 160, 12, 0.125, P, P, None) == -1
    assert lib.bvc_attn_fwd(P, 2, 0, 12, 0.125, P, P, None) == -1
    assert lib.bvc_attn_bwd(P, P, P + 2, P, 2, 160, 12, 0.125, P, P, None, 0, None) == -1
    assert lib.bvc_mask_to_index(P, 2, 1568, 1569, P, P, P, P, None) == -1
