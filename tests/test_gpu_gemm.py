"""tcgen05 3xTF32 GEMM (vkp_gemm) against float64 NumPy and against the SIMT fp32 kernel.

Tolerance: each fp32 operand is split into two TF32 numbers with round-to-nearest
(|x - hi - lo| <= 2^-22 |x|) and the lo*lo term is dropped (<= 2^-22 |a b|), so every product
is within ~3 * 2^-22 = 7.2e-7 of exact; the tensor core then adds the products of a k-step into
the fp32 TMEM accumulator with a truncating, exponent-aligned add (measured: up to ~4e-6 of
sum|a||b| at K = 8192).  The reference accumulates serially in fp32 (matmul.comp:29-31; bound
K * 2^-24 * sum|a||b|, i.e. 4.9e-4 at K = 8192).  Bound used here, element-wise:
|C - C_exact| <= TOL * (|A| |B|) with TOL = 6e-6."""
import numpy as np
import pytest

import vulkpy_b200 as vk

pytestmark = pytest.mark.gpu
F = np.float32
SIMT, TC, ACC = 1, 2, 4
TOL = 6e-6


def gemm(gpu, a, b, ta=False, tb=False, bias=None, flags=0, c0=None):
    A, B = vk.Array(gpu, data=a), vk.Array(gpu, data=b)
    M = a.shape[1] if ta else a.shape[0]
    K = a.shape[0] if ta else a.shape[1]
    N = b.shape[0] if tb else b.shape[1]
    C = vk.Array(gpu, data=c0) if c0 is not None else vk.Array(gpu, shape=(M, N))
    bb = vk.Array(gpu, data=bias) if bias is not None else None
    C.job = gpu.gpu.gemm(ta, tb, M, N, K, A.buffer, B.buffer, C.buffer, bb.buffer if bb is not None else None, flags)
    return np.asarray(C)


def exact(a, b, ta, tb):
    a64, b64 = a.astype(np.float64), b.astype(np.float64)
    a64 = a64.T if ta else a64
    b64 = b64.T if tb else b64
    return a64 @ b64, np.abs(a64) @ np.abs(b64)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 64), (256, 128, 96), (128, 256, 128), (384, 512, 160),
                                   (1024, 1024, 1024), (132, 260, 36), (1000, 136, 500), (4096, 256, 64)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
def test_tc_gemm_vs_float64(gpu, rs, M, N, K, ta, tb):
    a = rs.uniform(-1, 1, (K, M) if ta else (M, K)).astype(F)
    b = rs.uniform(-1, 1, (N, K) if tb else (K, N)).astype(F)
    got = gemm(gpu, a, b, ta, tb, flags=TC)
    want, mag = exact(a, b, ta, tb)
    err = np.abs(got - want) / mag
    assert err.max() < TOL, err.max()
    simt = gemm(gpu, a, b, ta, tb, flags=SIMT)
    assert (np.abs(simt - want) / mag).max() < 2e-6
    # the tensor core accumulates each k-step into TMEM with its own (truncating) fp32 add, so its
    # error exceeds the FFMA kernel's by a small factor while staying far inside the bound above
    assert np.abs(got - simt).max() < TOL * mag.max()


def test_tc_gemm_bias_and_accumulate(gpu, rs):
    M, N, K = 256, 384, 128
    a = rs.normal(size=(M, K)).astype(F)
    b = rs.normal(size=(N, K)).astype(F)
    bias = rs.normal(size=N).astype(F)
    c0 = rs.normal(size=(M, N)).astype(F)
    want, mag = exact(a, b, False, True)
    got = gemm(gpu, a, b, False, True, bias=bias, flags=TC)
    assert (np.abs(got - (want + bias)) / (mag + 1)).max() < TOL
    got = gemm(gpu, a, b, False, True, bias=bias, flags=TC | ACC, c0=c0)
    assert (np.abs(got - (want + bias + c0)) / (mag + 2)).max() < TOL
    got = gemm(gpu, a, b, False, True, flags=SIMT | ACC, c0=c0)
    assert (np.abs(got - (want + c0)) / (mag + 1)).max() < 2e-6


def test_tc_gemm_wide_dynamic_range(gpu, rs):
    """3xTF32 keeps fp32 accuracy when magnitudes vary by many orders (plain TF32 would lose 13 bits)."""
    M = N = K = 256
    a = (rs.normal(size=(M, K)) * np.exp2(rs.integers(-20, 20, (M, K)))).astype(F)
    b = (rs.normal(size=(N, K)) * np.exp2(rs.integers(-20, 20, (N, K)))).astype(F)
    got = gemm(gpu, a, b, False, True, flags=TC)
    want, mag = exact(a, b, False, True)
    assert (np.abs(got - want) / mag).max() < TOL
    sp = np.array([[np.inf, 1.0], [np.nan, 2.0]], dtype=F)
    big_a = np.zeros((128, 32), F); big_a[:2, :2] = sp
    big_b = np.zeros((128, 32), F); big_b[:, 0] = 1.0
    out = gemm(gpu, big_a, big_b, False, True, flags=TC)
    # non-finite inputs stay non-finite (inf * lo(=0) makes the split product NaN), finite rows are untouched
    assert not np.isfinite(out[0, 0]) and np.isnan(out[1, 0]) and out[2, 0] == 0


def test_matmul_and_dense_use_the_tensor_core_path(gpu, rs):
    """`@` and nn.Dense at tile-aligned sizes dispatch to the tcgen05 kernel and stay fp32-accurate."""
    from vulkpy_b200 import nn
    a = rs.uniform(-1, 1, (512, 256)).astype(F)
    b = rs.uniform(-1, 1, (256, 384)).astype(F)
    got = np.asarray(vk.Array(gpu, data=a) @ vk.Array(gpu, data=b))
    want, mag = exact(a, b, False, False)
    assert (np.abs(got - want) / mag).max() < TOL
    w = rs.normal(size=(256, 128)).astype(F)
    bias = rs.normal(size=256).astype(F)
    x = rs.normal(size=(512, 128)).astype(F)
    d = nn.Dense(gpu, 128, 256, w_init=lambda g, s: vk.Array(g, data=w), b_init=lambda g, s: vk.Array(g, data=bias))
    y = np.asarray(d(vk.Array(gpu, data=x)))
    want = x.astype(np.float64) @ w.astype(np.float64).T + bias
    assert np.abs(y - want).max() < 2e-4
    dy = rs.normal(size=(512, 256)).astype(F)
    dx = np.asarray(d.backward(vk.Array(gpu, data=dy)))
    assert np.abs(dx - dy.astype(np.float64) @ w).max() < 5e-4
    assert np.abs(np.asarray(d.w.grad) - dy.astype(np.float64).T @ x).max() < 1e-3


@pytest.mark.parametrize("M,N,K,ta,tb", [(16, 1024, 8192, True, False), (8192, 16, 1024, False, True),
                                         (8192, 1024, 16, False, False), (1024, 1024, 8192, True, False),
                                         (256, 512, 4096, False, True), (3, 5, 5000, False, False)])
def test_skinny_and_split_k(gpu, rs, M, N, K, ta, tb):
    """Shapes of the MLP step (C = 16 classes, batch 8192): few output tiles and a long K are split
    over K (SIMT: blockIdx.z slices, tcgen05: (tile, split) work items) and reduced in a fixed order."""
    a = rs.uniform(-1, 1, (K, M) if ta else (M, K)).astype(F)
    b = rs.uniform(-1, 1, (N, K) if tb else (K, N)).astype(F)
    bias = rs.normal(size=N).astype(F)
    c0 = rs.normal(size=(M, N)).astype(F)
    want, mag = exact(a, b, ta, tb)
    got = gemm(gpu, a, b, ta, tb)
    assert (np.abs(got - want) / mag).max() < TOL
    got = gemm(gpu, a, b, ta, tb, bias=bias, flags=ACC, c0=c0)
    assert (np.abs(got - (want + bias + c0)) / (mag + 2)).max() < TOL
    got2 = gemm(gpu, a, b, ta, tb, bias=bias, flags=ACC, c0=c0)
    np.testing.assert_array_equal(got, got2)          # deterministic


@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
def test_tc_gemm_presplit_path(gpu, rs, ta, tb):
    """M, N >= 2048 takes the pre-split variant (lo tiles produced once in HBM, fused with the
    transpose; 4-stage BK=16 ring): several k-block ring wraps per tile, same bound."""
    M, N, K = 2048, 2304, 272
    a = rs.uniform(-1, 1, (K, M) if ta else (M, K)).astype(F)
    b = rs.uniform(-1, 1, (N, K) if tb else (K, N)).astype(F)
    got = gemm(gpu, a, b, ta, tb, flags=TC)
    want, mag = exact(a, b, ta, tb)
    assert (np.abs(got - want) / mag).max() < TOL
    bias = rs.normal(size=N).astype(F)
    c0 = rs.normal(size=(M, N)).astype(F)
    got = gemm(gpu, a, b, ta, tb, bias=bias, flags=TC | ACC, c0=c0)
    assert (np.abs(got - (want + bias + c0)) / (mag + 2)).max() < TOL


@pytest.mark.parametrize("case", ["n16", "n10", "m16", "m7", "k16", "k5"])
def test_skinny_gemm_kernels(gpu, rs, case):
    """The three streaming kernels behind nn.Dense with few classes (forward N <= 16, weight gradient
    M <= 16, input gradient K <= 16; vkp_gemm.cu): float32 FMA accumulation, bound 2e-6 * |A||B|, with
    bias / accumulate where the callers use them."""
    if case[0] == "n":      # Y = X W^T + b   (batch_affine.comp:25-39)
        N = int(case[1:]); M, K, ta, tb = 2048, 1024, False, True
    elif case[0] == "m":    # dW = dy^T x     (nn/layers.py:126-141 as one contraction)
        M = int(case[1:]); N, K, ta, tb = 1024, 4096, True, False
    else:                   # dx = dy W
        K = int(case[1:]); M, N, ta, tb = 2048, 1024, False, False
    a = rs.uniform(-1, 1, (K, M) if ta else (M, K)).astype(F)
    b = rs.uniform(-1, 1, (N, K) if tb else (K, N)).astype(F)
    want, mag = exact(a, b, ta, tb)
    # N = 16 (a multiple of 4) forward takes the narrow tcgen05 tile (128 x 32, 3xTF32 bound); the others stream in FFMA
    tol = TOL if case == "n16" else 2e-6
    got = gemm(gpu, a, b, ta, tb)
    assert (np.abs(got - want) / (mag + 1e-30)).max() < tol
    c0 = rs.normal(size=(M, N)).astype(F)
    bias = rs.normal(size=N).astype(F) if case[0] != "k" else None
    got = gemm(gpu, a, b, ta, tb, bias=bias, flags=ACC, c0=c0)
    ref = want + c0 + (bias if bias is not None else 0)
    assert (np.abs(got - ref) / (mag + 2)).max() < tol
    simt = gemm(gpu, a, b, ta, tb, flags=SIMT)
    assert (np.abs(got - (simt + c0 + (bias if bias is not None else 0))) / (mag + 2)).max() < tol


def test_fused_relu_epilogues_equal_separate_kernels(gpu, rs):
    """VKP_GEMM_RELU (Dense + ReLU forward) and the relu_mask epilogue (ReLU.backward on the dx GEMM) are the
    stand-alone kernels' operations on the same GEMM result: bit-identical, on the tcgen05, skinny and SIMT paths."""
    RELU = 8
    for (M, N, K, ta, tb) in ((512, 256, 128, False, True), (8192, 1024, 16, False, False), (2048, 16, 1024, False, True),
                              (96, 80, 72, False, False), (1024, 1024, 1024, False, True)):
        a = rs.normal(size=(K, M) if ta else (M, K)).astype(F)
        b = rs.normal(size=(N, K) if tb else (K, N)).astype(F)
        bias = rs.normal(size=N).astype(F)
        y = rs.normal(size=(M, N)).astype(F)
        y[0, :4] = [0.0, -0.0, 1.0, -1.0]
        plain = gemm(gpu, a, b, ta, tb, bias=bias)
        fused = gemm(gpu, a, b, ta, tb, bias=bias, flags=RELU)
        np.testing.assert_array_equal(fused.view(np.uint32), np.asarray(vk.Array(gpu, data=plain).max(0.0)).view(np.uint32))
        A, B, Y = vk.Array(gpu, data=a), vk.Array(gpu, data=b), vk.Array(gpu, data=y)
        C = vk.Array(gpu, shape=(M, N))
        C.job = gpu.gpu.gemm(ta, tb, M, N, K, A.buffer, B.buffer, C.buffer, None, 0, relu_mask=Y.buffer)
        nob = vk.Array(gpu, data=gemm(gpu, a, b, ta, tb))
        sep = vk.Array(gpu, shape=(M, N))
        sep.job = gpu.gpu.nn_activation_backward(0, Y.buffer, nob.buffer, sep.buffer)
        np.testing.assert_array_equal(np.asarray(C).view(np.uint32), np.asarray(sep).view(np.uint32))


@pytest.mark.parametrize("M,N,K,tb", [(2048, 8192, 512, True), (2048, 8192 + 64, 256, True), (1024, 8192, 256, True)])
def test_tc_bn224_tile(gpu, M, N, K, tb):
    """Shapes whose 128x256 tiling wastes most of a wave take 128x224 tiles (vkp_gemm_tc.cu prefer_bn224): 37 column
    tiles of which the last is ragged (TMA zero fill + masked epilogue), TMEM allocation rounded up to 512 columns,
    UMMA N = 224.  K-major B (the pre-split NT form Dense.forward and the row-sharded matmul use), bias + ReLU epilogue."""
    rs = np.random.default_rng(M + N)
    a = rs.uniform(-1, 1, (M, K)).astype(F)
    b = rs.uniform(-1, 1, (N, K) if tb else (K, N)).astype(F)
    want, mag = exact(a, b, False, tb)
    got = gemm(gpu, a, b, tb=tb, flags=TC)
    assert (np.abs(got - want) / mag).max() < TOL
    bias = rs.uniform(-1, 1, N).astype(F)
    got = gemm(gpu, a, b, tb=tb, bias=bias, flags=TC)
    assert (np.abs(got - (want + bias)) / (mag + np.abs(bias))).max() < TOL
