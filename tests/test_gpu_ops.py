"""-m gpu: per-op parity of the CUDA kernels against the oracle / golden fixtures, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import bigvgan_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_activation1d_golden_fp32(golden_dir, case):
    from tests import gpu_util as G
    g = np.load(os.path.join(golden_dir, "activation1d.npz"))
    y = G.activation1d(g[f"{case}_x"], g[f"{case}_alpha"], g[f"{case}_beta"])
    np.testing.assert_allclose(y, g[f"{case}_y"], atol=1e-5)


@pytest.mark.parametrize("shape", [(2, 40, 1000), (1, 3, 5), (3, 8, 257), (1, 24, 4096 + 17)])
def test_activation1d_oracle_fp32(shape):
    from tests import gpu_util as G
    rng = np.random.default_rng(hash(shape) % 1000)
    x = (1.5 * rng.standard_normal(shape)).astype(np.float32)
    la = (0.5 * rng.standard_normal(shape[1])).astype(np.float32)
    lb = (0.5 * rng.standard_normal(shape[1])).astype(np.float32)
    ref = O.activation1d(x.astype(np.float64), la.astype(np.float64), lb.astype(np.float64))
    np.testing.assert_allclose(G.activation1d(x, la, lb), ref, atol=2e-5)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 4e-2), (torch.float16, 6e-3)])
def test_activation1d_half_types(dtype, tol):
    from tests import gpu_util as G
    rng = np.random.default_rng(5)
    x = (1.5 * rng.standard_normal((2, 16, 700))).astype(np.float32)
    la = (0.5 * rng.standard_normal(16)).astype(np.float32)
    lb = (0.5 * rng.standard_normal(16)).astype(np.float32)
    xr = torch.as_tensor(x).to(dtype).float().numpy().astype(np.float64)
    ref = O.activation1d(xr, la.astype(np.float64), lb.astype(np.float64))
    y = G.activation1d(x, la, lb, dtype)
    assert np.abs(y - ref).max() <= tol * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("shape", [(1, 8, 1), (2, 24, 7), (1, 8, 35), (2, 16, 131), (1, 48, 256), (3, 8, 1000),
                                   (1, 24, 2 * 4096 + 301),
                                   # tile-boundary cases of the tensor-core kernel: a tile ending 0..3 rows before the segment end
                                   (1, 8, 2), (1, 8, 3), (1, 16, 8), (1, 8, 9), (2, 8, 223), (1, 16, 224), (1, 8, 225), (1, 8, 226),
                                   (1, 8, 227), (1, 24, 449), (1, 8, 2 * 224 + 2), (16, 768, 96)])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_activation1d_packed_kernel(shape, mode):
    """The hot-path kernel (packed c8 layout): interior fast path, sequence ends, short segments.  Mode 2 (fp16
    storage) is the sharp check of the tensor-core kernel's segment-end handling: 3e-3 instead of bf16's 2e-2."""
    from tests import gpu_util as G
    rng = np.random.default_rng(sum(shape))
    x = (1.5 * rng.standard_normal(shape)).astype(np.float32)
    la = (0.5 * rng.standard_normal(shape[1])).astype(np.float32)
    lb = (0.5 * rng.standard_normal(shape[1])).astype(np.float32)
    xin = x.astype(np.float64) if mode == 0 else (G.bf16_round(x) if mode == 1 else G.f16_round(x))
    ref = O.activation1d(xin, la.astype(np.float64), lb.astype(np.float64))
    y = G.activation1d_packed(x, la, lb, mode)
    tol = 2e-5 if mode == 0 else (2e-2 if mode == 1 else 3e-3) * max(1.0, np.abs(ref).max())
    assert np.abs(y - ref).max() <= tol, np.abs(y - ref).max()


def test_activation1d_empty_is_noop():
    from tests import gpu_util as G
    from b200vgan import lib
    z = torch.zeros(1, device="cuda")
    lib.check(G.L_().bvg_activation1d(z.data_ptr(), z.data_ptr(), z.data_ptr(), z.data_ptr(), 1, 8, 0, 0, None))


CONV_CASES = [  # (B, Cin, Cout, T, k, d)
    (1, 16, 24, 50, 3, 1), (2, 24, 24, 300, 7, 3), (1, 64, 48, 257, 11, 5), (2, 8, 8, 1, 3, 1),
    (1, 96, 96, 200, 7, 1), (1, 1024, 64, 40, 7, 1),
]


def _conv_inputs(B, Cin, Cout, T, k, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, k)) / np.sqrt(Cin * k)).astype(np.float32)
    b = (0.1 * rng.standard_normal(Cout)).astype(np.float32)
    r = rng.standard_normal((B, Cout, T)).astype(np.float32)
    return x, w, b, r


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv1d_cuda_core_fp32(case):
    from tests import gpu_util as G
    B, Cin, Cout, T, k, d = case
    x, w, b, r = _conv_inputs(B, Cin, Cout, T, k, 1)
    ref = O.conv1d(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), dilation=d,
                   padding=O.get_padding(k, d)) + r
    y = G.conv1d(x, w, b, r, k, d, 0)
    np.testing.assert_allclose(y, ref, atol=2e-5, rtol=1e-5)
    y = G.conv1d(x, w, None, None, k, d, 0)
    np.testing.assert_allclose(y, ref - r - b[None, :, None], atol=2e-5, rtol=1e-5)


# the last three have p = (k-u)/2 > u: the final rows of a segment need q beyond len_in + 1 (q_extra = ceil(p/u))
CONVT_CASES = [(1, 32, 16, 37, 8, 4), (2, 16, 8, 5, 4, 4), (1, 48, 24, 130, 4, 2), (1, 64, 32, 1, 8, 4),
               (2, 16, 8, 131, 16, 4), (1, 24, 16, 3, 12, 2), (1, 8, 8, 128, 20, 4)]


@pytest.mark.parametrize("case", CONVT_CASES)
def test_conv_transpose1d_cuda_core_fp32(case):
    from tests import gpu_util as G
    B, Cin, Cout, T, k, u = case
    rng = np.random.default_rng(2)
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cin, Cout, k)) / np.sqrt(Cin * k / u)).astype(np.float32)
    b = (0.1 * rng.standard_normal(Cout)).astype(np.float32)
    ref = O.conv_transpose1d(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), u, (k - u) // 2)
    np.testing.assert_allclose(G.conv_transpose1d(x, w, b, k, u, 0), ref, atol=2e-5, rtol=1e-5)


UMMA_CONV_CASES = [  # (B, Cin, Cout, T, k, d)
    (1, 64, 64, 128, 3, 1), (1, 64, 64, 300, 3, 1), (2, 192, 192, 260, 7, 3), (1, 96, 96, 515, 11, 5),
    (1, 48, 48, 700, 7, 5), (1, 24, 24, 1000, 11, 1), (1, 128, 128, 130, 3, 3), (1, 256, 768, 300, 3, 5),
    (2, 384, 384, 200, 11, 5), (1, 1024, 1536, 100, 7, 1), (3, 768, 768, 77, 7, 1),
    # more 512-row tiles than SMs: persistent CTAs walk several tiles (stage rings wrap); K-packed 24-channel layers, odd / even taps
    (1, 24, 24, 100000, 11, 5), (1, 24, 24, 90000, 7, 3), (2, 24, 24, 45000, 3, 1), (1, 48, 48, 90000, 3, 1),
]


@pytest.mark.parametrize("case", UMMA_CONV_CASES)
def test_conv1d_tcgen05_bf16(case):
    from tests import gpu_util as G
    B, Cin, Cout, T, k, d = case
    x, w, b, r = _conv_inputs(B, Cin, Cout, T, k, 3)
    ref = O.conv1d(G.bf16_round(x), G.bf16_round(w), b.astype(np.float64), dilation=d,
                   padding=O.get_padding(k, d)) + G.bf16_round(r)
    y = G.conv1d(x, w, b, r, k, d, 1)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() <= 8e-3 * scale, (np.abs(y - ref).max(), scale)


@pytest.mark.parametrize("case", UMMA_CONV_CASES)
def test_conv1d_tcgen05_fp16(case):
    """fp16 storage mode (BVG_MODE_F16 = 2): kind::f16 MMAs on fp16 operands, saturating fp16 stores."""
    from tests import gpu_util as G
    B, Cin, Cout, T, k, d = case
    x, w, b, r = _conv_inputs(B, Cin, Cout, T, k, 3)
    ref = O.conv1d(G.f16_round(x), G.f16_round(w), b.astype(np.float64), dilation=d,
                   padding=O.get_padding(k, d)) + G.f16_round(r)
    y = G.conv1d(x, w, b, r, k, d, 2)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() <= 1e-3 * scale, (np.abs(y - ref).max(), scale)


@pytest.mark.parametrize("case", [(1, 64, 32, 100, 8, 4), (1, 48, 24, 300, 4, 2), (2, 64, 32, 127, 16, 4)])
def test_conv_transpose1d_tcgen05_fp16(case):
    from tests import gpu_util as G
    B, Cin, Cout, T, k, u = case
    rng = np.random.default_rng(4)
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cin, Cout, k)) / np.sqrt(Cin * k / u)).astype(np.float32)
    b = (0.1 * rng.standard_normal(Cout)).astype(np.float32)
    ref = O.conv_transpose1d(G.f16_round(x), G.f16_round(w), b.astype(np.float64), u, (k - u) // 2)
    y = G.conv_transpose1d(x, w, b, k, u, 2)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() <= 1e-3 * scale, (np.abs(y - ref).max(), scale)


@pytest.mark.parametrize("case", [(1, 64, 32, 100, 8, 4), (2, 192, 96, 50, 4, 4), (1, 48, 24, 300, 4, 2),
                                  (1, 1536, 768, 40, 8, 4), (1, 96, 48, 129, 4, 2),
                                  (2, 64, 32, 127, 16, 4), (1, 48, 24, 256, 12, 2)])   # p > u (k = 4u, 6u)
def test_conv_transpose1d_tcgen05_bf16(case):
    from tests import gpu_util as G
    B, Cin, Cout, T, k, u = case
    rng = np.random.default_rng(4)
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cin, Cout, k)) / np.sqrt(Cin * k / u)).astype(np.float32)
    b = (0.1 * rng.standard_normal(Cout)).astype(np.float32)
    ref = O.conv_transpose1d(G.bf16_round(x), G.bf16_round(w), b.astype(np.float64), u, (k - u) // 2)
    y = G.conv_transpose1d(x, w, b, k, u, 1)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() <= 8e-3 * scale, (np.abs(y - ref).max(), scale)


ACT_CONV_CASES = [  # (B, C, T, k, d): AMPBlock1 half-steps (Activation1d -> Conv1d [+ residual])
    (1, 24, 700, 3, 1), (1, 24, 100, 11, 5), (2, 48, 333, 7, 3), (2, 96, 300, 11, 5), (1, 192, 400, 3, 1),
    (3, 24, 7, 3, 1), (1, 384, 90, 7, 1),
    (1, 24, 80000, 11, 1),   # several tiles per persistent CTA, with and without the residual identity stages
]


@pytest.mark.parametrize("case", ACT_CONV_CASES)
@pytest.mark.parametrize("want_fused", [False, True])
@pytest.mark.parametrize("with_res", [False, True])
def test_act_conv1d_tcgen05_bf16(case, want_fused, with_res):
    """Activation1d + conv through both bf16 paths: separate passes (what bvg_forward runs) and the
    experimental producer-fused kernel; layers the fused kernel does not cover must fall back."""
    from tests import gpu_util as G
    B, C, T, k, d = case
    rng = np.random.default_rng(11)
    x = rng.standard_normal((B, C, T)).astype(np.float32)
    la = (0.3 * rng.standard_normal(C)).astype(np.float32)
    lb = (0.3 * rng.standard_normal(C)).astype(np.float32)
    w = (rng.standard_normal((C, C, k)) / np.sqrt(C * k)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    r = rng.standard_normal((B, C, T)).astype(np.float32) if with_res else None
    act = O.activation1d(G.bf16_round(x), la.astype(np.float64), lb.astype(np.float64))
    ref = O.conv1d(G.bf16_round(act), G.bf16_round(w), b.astype(np.float64), dilation=d, padding=O.get_padding(k, d))
    if with_res:
        ref = ref + G.bf16_round(r)
    y, fused = G.act_conv1d(x, la, lb, w, b, r, k, d, 1, want_fused=want_fused)
    assert fused == (want_fused and C <= 256)
    scale = np.abs(ref).max()
    # the activated tensor is rounded to bf16 before the MMA on both paths; allow one more rounding step
    assert np.abs(y - ref).max() <= 1.6e-2 * scale, (np.abs(y - ref).max(), scale)


def test_act_conv1d_cuda_core_fp32():
    from tests import gpu_util as G
    B, C, T, k, d = 2, 24, 211, 7, 3
    rng = np.random.default_rng(12)
    x = rng.standard_normal((B, C, T)).astype(np.float32)
    la = (0.3 * rng.standard_normal(C)).astype(np.float32)
    lb = (0.3 * rng.standard_normal(C)).astype(np.float32)
    w = (rng.standard_normal((C, C, k)) / np.sqrt(C * k)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    ref = O.conv1d(O.activation1d(x.astype(np.float64), la.astype(np.float64), lb.astype(np.float64)),
                   w.astype(np.float64), b.astype(np.float64), dilation=d, padding=O.get_padding(k, d))
    y, fused = G.act_conv1d(x, la, lb, w, b, None, k, d, 0, want_fused=True)
    assert not fused
    np.testing.assert_allclose(y, ref, atol=3e-5, rtol=1e-5)


# ---- fp32 tensor-core mode (BVG_MODE_FP32_TC = 3): fp32 storage, three bf16 tcgen05 products per convolution ----
TC32_CONV_CASES = [(1, 64, 64, 300, 3, 1), (2, 192, 192, 260, 7, 3), (1, 96, 96, 515, 11, 5), (1, 48, 48, 700, 7, 5),
                   (1, 24, 24, 1000, 11, 1), (1, 256, 768, 300, 3, 5), (1, 1024, 1536, 100, 7, 1), (3, 768, 768, 77, 7, 1),
                   (1, 24, 24, 90000, 7, 3)]


@pytest.mark.parametrize("case", TC32_CONV_CASES)
def test_conv1d_tcgen05_fp32_split(case):
    """x_hi w_hi + x_lo w_hi + x_hi w_lo with fp32 accumulation against the fp64 oracle on the UNROUNDED fp32 inputs."""
    from tests import gpu_util as G
    B, Cin, Cout, T, k, d = case
    x, w, b, r = _conv_inputs(B, Cin, Cout, T, k, 3)
    ref = O.conv1d(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), dilation=d,
                   padding=O.get_padding(k, d)) + r
    y = G.conv1d(x, w, b, r, k, d, 3)
    scale = np.abs(ref).max()
    err = np.abs(y - ref).max()
    print("fp32-tc conv", case, "max-abs", err, "scale", scale)
    assert err <= 5e-5 * scale, (err, scale)


@pytest.mark.parametrize("case", [(1, 64, 32, 100, 8, 4), (1, 48, 24, 300, 4, 2), (2, 64, 32, 127, 16, 4), (1, 1536, 768, 40, 8, 4)])
def test_conv_transpose1d_tcgen05_fp32_split(case):
    from tests import gpu_util as G
    B, Cin, Cout, T, k, u = case
    rng = np.random.default_rng(4)
    x = rng.standard_normal((B, Cin, T)).astype(np.float32)
    w = (rng.standard_normal((Cin, Cout, k)) / np.sqrt(Cin * k / u)).astype(np.float32)
    b = (0.1 * rng.standard_normal(Cout)).astype(np.float32)
    ref = O.conv_transpose1d(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), u, (k - u) // 2)
    y = G.conv_transpose1d(x, w, b, k, u, 3)
    scale = np.abs(ref).max()
    err = np.abs(y - ref).max()
    print("fp32-tc conv-transpose", case, "max-abs", err, "scale", scale)
    assert err <= 5e-5 * scale, (err, scale)


@pytest.mark.parametrize("case", [(1, 24, 700, 3, 1), (2, 96, 300, 11, 5), (1, 192, 400, 3, 1)])
def test_act_conv1d_tcgen05_fp32_split(case):
    from tests import gpu_util as G
    B, C, T, k, d = case
    rng = np.random.default_rng(11)
    x = rng.standard_normal((B, C, T)).astype(np.float32)
    la = (0.3 * rng.standard_normal(C)).astype(np.float32)
    lb = (0.3 * rng.standard_normal(C)).astype(np.float32)
    w = (rng.standard_normal((C, C, k)) / np.sqrt(C * k)).astype(np.float32)
    b = (0.1 * rng.standard_normal(C)).astype(np.float32)
    r = rng.standard_normal((B, C, T)).astype(np.float32)
    ref = O.conv1d(O.activation1d(x.astype(np.float64), la.astype(np.float64), lb.astype(np.float64)),
                   w.astype(np.float64), b.astype(np.float64), dilation=d, padding=O.get_padding(k, d)) + r
    y, fused = G.act_conv1d(x, la, lb, w, b, r, k, d, 3)
    assert not fused
    err = np.abs(y - ref).max()
    print("fp32-tc act+conv", case, "max-abs", err, "scale", np.abs(ref).max())
    assert err <= 1e-4 * np.abs(ref).max()
