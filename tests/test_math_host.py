"""vkp_math.cuh compiled for the HOST (it is __host__ __device__): exp/exp2/log/log2/pow are the
same code the kernels run, so their accuracy is verified here, on the CPU, against float64."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = """
#include "vkp_math.cuh"
extern "C" {
void h_exp(const float* x, float* y, long n){ for(long i=0;i<n;i++) y[i]=vkpm::exp_f(x[i]); }
void h_exp2(const float* x, float* y, long n){ for(long i=0;i<n;i++) y[i]=vkpm::exp2_f(x[i]); }
void h_log(const float* x, float* y, long n){ for(long i=0;i<n;i++) y[i]=vkpm::log_f(x[i]); }
void h_log2(const float* x, float* y, long n){ for(long i=0;i<n;i++) y[i]=vkpm::log2_f(x[i]); }
void h_pow(const float* x, const float* y, float* z, long n){ for(long i=0;i<n;i++) z[i]=vkpm::pow_f(x[i],y[i]); }
// table-driven fast paths (what the kernels run for ordinary inputs), host table accessor
void f_exp(const float* x, float* y, long n){ vkpm::HostTables t; for(long i=0;i<n;i++) y[i]=vkpm::exp_fast(x[i],t); }
void f_exp2(const float* x, float* y, long n){ vkpm::HostTables t; for(long i=0;i<n;i++) y[i]=vkpm::exp2_fast(x[i],t); }
void f_log(const float* x, float* y, long n){ vkpm::HostTables t; for(long i=0;i<n;i++) y[i]=vkpm::log_fast(x[i],t); }
void f_log2(const float* x, float* y, long n){ vkpm::HostTables t; for(long i=0;i<n;i++) y[i]=vkpm::log2_fast(x[i],t); }
void f_pow(const float* x, const float* y, float* z, long n){ vkpm::HostTables t; for(long i=0;i<n;i++) z[i]=vkpm::pow_fast(x[i],y[i],t); }
int p_plan(float s){ vkpm::PowsCoef c; return vkpm::pows_plan(s, c); }
void p_pows(const float* x, float s, float* z, long n){
  vkpm::PowsCoef c; const int D = vkpm::pows_plan(s, c);
  vkpm::PowsHostTables t(s, c);
  for(long i=0;i<n;i++){
    const uint32_t u = vkpm::f2bits(x[i]);
    if (D == 0 || !((u - 0x00800000u) < 0x7f000000u)) { z[i] = vkpm::pow_f(x[i], s); continue; }
    z[i] = D == 6 ? vkpm::pows_core<6>(u, t) : D == 8 ? vkpm::pows_core<8>(u, t) : vkpm::pows_core<10>(u, t);
  }
}
void f_sincos(const float* x, float* s, float* c, long n){ for(long i=0;i<n;i++) vkpm::sincos_small(x[i], s[i], c[i]); }
void f_asinh(const float* x, float* y, long n){ for(long i=0;i<n;i++) y[i]=vkpm::asinh_f(x[i]); }
void f_bm_log(const float* x, float* y, long n){ for(long i=0;i<n;i++) y[i]=vkpm::bm_log(x[i]); }
void f_bm_pair(const float* u0, const float* u1, float* o, long n, float mean, float sd){
  for(long i=0;i<n;i++) vkpm::box_muller_pair(u0[i], u1[i], mean, sd, o[2*i], o[2*i+1]); }
}
"""


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    d = tmp_path_factory.mktemp("hostmath")
    src = d / "h.cpp"
    src.write_text(SRC)
    so = d / "h.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I",
                    os.path.join(ROOT, "vulkpy_b200", "csrc"), "-o", str(so), str(src)], check=True)
    return C.CDLL(str(so))


def call1(lib, name, x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    getattr(lib, name)(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_long(x.size))
    return y


def call2(lib, x, y, name="h_pow"):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    z = np.empty_like(x)
    getattr(lib, name)(C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_void_p(z.ctypes.data), C.c_long(x.size))
    return z


def ulps(got, exact):
    with np.errstate(all="ignore"):
        r32 = exact.astype(np.float32)
    return np.abs(got.astype(np.float64) - exact) / np.spacing(np.abs(r32)).astype(np.float64)


def bits(v):
    return [hex(int(b)) for b in np.asarray(v, dtype=np.float32).view(np.uint32)]


def test_appendix_a_points(hm):
    """correctly rounded results the reference's rtol=1e-7 tests need."""
    x = [1, 2, 3]
    assert bits(call1(hm, "h_exp", x)) == ["0x402df854", "0x40ec7326", "0x41a0af2e"]
    assert bits(call1(hm, "h_log", x)) == ["0x0", "0x3f317218", "0x3f8c9f54"]
    assert list(call1(hm, "h_exp2", x)) == [2.0, 4.0, 8.0]
    assert bits(call1(hm, "h_log2", x)) == ["0x0", "0x3f800000", "0x3fcae00d"]
    assert bits(call2(hm, x, [1.1, 2.2, 1.4])) == ["0x3f800000", "0x4093088d", "0x4094fa28"]
    assert bits(call2(hm, x, [2.7] * 3)) == ["0x3f800000", "0x40cfefc6", "0x419b5a2a"]
    assert bits(call2(hm, [1.3] * 3, [1.1, 2.2, 1.4])) == ["0x3faad2d2", "0x3fe3f958", "0x3fb8cfec"]
    assert list(call2(hm, [1, 2, 3, 4], [2, 3, 2, 3])) == [1.0, 8.0, 9.0, 64.0]


@pytest.mark.parametrize("name,f,lo,hi", [("h_exp", np.exp, -87.0, 88.0), ("h_exp2", np.exp2, -126.0, 127.0)])
def test_exp_accuracy(hm, name, f, lo, hi):
    x = np.random.default_rng(0).uniform(lo, hi, 2_000_000).astype(np.float32)
    got = call1(hm, name, x)
    assert ulps(got, f(x.astype(np.float64))).max() < 0.5002


@pytest.mark.parametrize("name,f", [("h_log", np.log), ("h_log2", np.log2)])
def test_log_accuracy(hm, name, f):
    rs = np.random.default_rng(1)
    x = rs.integers(1, 0x7f800000, 2_000_000, dtype=np.uint32).view(np.float32)  # every positive finite float
    assert ulps(call1(hm, name, x), f(x.astype(np.float64))).max() < 0.5002
    x = rs.uniform(0.5, 2.0, 1_000_000).astype(np.float32)
    e = ulps(call1(hm, name, x), f(x.astype(np.float64)))
    assert e[np.isfinite(e)].max() < 0.5002


def test_pow_accuracy(hm):
    rs = np.random.default_rng(2)
    x = rs.uniform(0.01, 100, 2_000_000).astype(np.float32)
    y = rs.uniform(-10, 10, 2_000_000).astype(np.float32)
    exact = np.power(x.astype(np.float64), y.astype(np.float64))
    with np.errstate(all="ignore"):
        ok = np.isfinite(exact.astype(np.float32)) & (exact.astype(np.float32) != 0)
    assert ulps(call2(hm, x, y)[ok], exact[ok]).max() < 0.5002


def test_special_values(hm):
    inf, nan = np.inf, np.nan
    assert list(call1(hm, "h_exp", [-inf, inf, -200, 200])) == [0.0, inf, 0.0, inf]
    assert np.isnan(call1(hm, "h_exp", [nan])[0])
    r = call1(hm, "h_log", [0.0, -1.0, inf, 1e-45])
    assert r[0] == -inf and np.isnan(r[1]) and r[2] == inf and abs(r[3] - np.log(1.401298464e-45)) < 1e-4
    # C99 pow corner cases
    got = call2(hm, [2, -2, -2, 0, -0.0, inf, -8, 2, 0.5, 7, 1, -1],
                    [0, 3, 2, -1, -3, -1, 1 / 3, inf, inf, -inf, nan, inf])
    want = [1, -8, 4, inf, -inf, 0, nan, inf, 0, 0, 1, 1]
    for g, w in zip(got, want):
        assert (np.isnan(g) and np.isnan(w)) or g == w, (got, want)
    # subnormal results are flushed to zero (test/test_nn.py:120-126 needs exp(-100) == 0), the
    # smallest normal survives
    x = np.float32([-140.5, -149.0, -126.0, -126.5])
    np.testing.assert_array_equal(call1(hm, "h_exp2", x), np.float32([0, 0, 2.0 ** -126, 0]))
    np.testing.assert_array_equal(call1(hm, "f_exp2", x), np.float32([0, 0, 2.0 ** -126, 0]))
    assert call1(hm, "h_exp", [-100.0])[0] == 0 and call1(hm, "f_exp", [-100.0])[0] == 0
    assert call2(hm, [1e-20], [2.5])[0] == 0 and call2(hm, [1e-20], [2.5], "f_pow")[0] == 0


# ---- the table-driven fast paths: same accuracy, same special-value behaviour ---------------------
def test_fast_paths_appendix_a(hm):
    x = [1, 2, 3]
    assert bits(call1(hm, "f_exp", x)) == ["0x402df854", "0x40ec7326", "0x41a0af2e"]
    assert bits(call1(hm, "f_log", x)) == ["0x0", "0x3f317218", "0x3f8c9f54"]
    assert list(call1(hm, "f_exp2", x)) == [2.0, 4.0, 8.0]
    assert bits(call1(hm, "f_log2", x)) == ["0x0", "0x3f800000", "0x3fcae00d"]
    assert bits(call2(hm, x, [1.1, 2.2, 1.4], "f_pow")) == ["0x3f800000", "0x4093088d", "0x4094fa28"]
    assert bits(call2(hm, x, [2.7] * 3, "f_pow")) == ["0x3f800000", "0x40cfefc6", "0x419b5a2a"]
    assert bits(call2(hm, [1.3] * 3, [1.1, 2.2, 1.4], "f_pow")) == ["0x3faad2d2", "0x3fe3f958", "0x3fb8cfec"]
    assert list(call2(hm, [1, 2, 3, 4], [2, 3, 2, 3], "f_pow")) == [1.0, 8.0, 9.0, 64.0]


def test_fast_paths_accuracy(hm):
    rs = np.random.default_rng(5)
    for name, f, lo, hi in [("f_exp", np.exp, -104.0, 89.0), ("f_exp2", np.exp2, -150.0, 128.0)]:
        x = rs.uniform(lo, hi, 2_000_000).astype(np.float32)
        exact = f(x.astype(np.float64))
        with np.errstate(all="ignore"):
            ok = np.isfinite(exact.astype(np.float32)) & (np.abs(exact) > 1.2e-38)
        assert ulps(call1(hm, name, x)[ok], exact[ok]).max() < 0.5002
        assert (call1(hm, name, x)[np.abs(exact) < 1.1e-38] == 0).all()
    for name, f in [("f_log", np.log), ("f_log2", np.log2)]:
        x = rs.integers(0x00800000, 0x7f800000, 2_000_000, dtype=np.uint32).view(np.float32)
        assert ulps(call1(hm, name, x), f(x.astype(np.float64))).max() < 0.5002
        x = (1 + rs.uniform(-2.1e-2, 2.1e-2, 1_000_000)).astype(np.float32)    # the interval that contains 1.0
        e = ulps(call1(hm, name, x), f(x.astype(np.float64)))
        assert e[np.isfinite(e)].max() < 0.5002
    x = rs.integers(1, 0x7f800000, 2_000_000, dtype=np.uint32).view(np.float32)
    y = rs.uniform(-3, 3, 2_000_000).astype(np.float32)
    exact = np.power(x.astype(np.float64), y.astype(np.float64))
    with np.errstate(all="ignore"):
        ok = np.isfinite(exact.astype(np.float32)) & (np.abs(exact) > 1.2e-38)
    assert ulps(call2(hm, x, y, "f_pow")[ok], exact[ok]).max() < 0.5002
    # fast and careful routines agree bit for bit except within 1e-4 ulp of a rounding boundary
    assert (call2(hm, x, y, "f_pow") != call2(hm, x, y)).mean() < 1e-4


def test_fast_paths_special_values(hm):
    inf, nan = np.inf, np.nan
    np.testing.assert_array_equal(call1(hm, "f_exp", [-inf, inf, -200, 200, 88.5, -100]),
                                  call1(hm, "h_exp", [-inf, inf, -200, 200, 88.5, -100]))
    r = call1(hm, "f_log", [0.0, -1.0, inf, 1e-45, nan])    # subnormal INPUTS are still honoured
    assert r[0] == -inf and np.isnan(r[1]) and r[2] == inf and abs(r[3] + 103.2789) < 1e-3 and np.isnan(r[4])
    xs = [2, -2, -2, 0, -0.0, inf, -8, 2, 0.5, 7, 1, -1, 1, 1e-40]
    ys = [0, 3, 2, -1, -3, -1, 1 / 3, inf, inf, -inf, nan, inf, 1e30, 2]
    a, b = call2(hm, xs, ys, "f_pow"), call2(hm, xs, ys)
    assert all((np.isnan(p) and np.isnan(q)) or p == q for p, q in zip(a, b))


def test_sincos_small(hm):
    u = np.random.default_rng(6).uniform(0, 1, 2_000_000).astype(np.float32)
    x = (np.float32(6.28318530718) * u).astype(np.float32)          # the Box-Muller angle
    s, c = np.empty_like(x), np.empty_like(x)
    hm.f_sincos(C.c_void_p(x.ctypes.data), C.c_void_p(s.ctypes.data), C.c_void_p(c.ctypes.data), C.c_long(x.size))
    assert np.abs(s - np.sin(x.astype(np.float64))).max() < 1.2e-7
    assert np.abs(c - np.cos(x.astype(np.float64))).max() < 1.2e-7
    assert ulps(s, np.sin(x.astype(np.float64))).max() < 2.0 and ulps(c, np.cos(x.astype(np.float64))).max() < 2.0


def test_bm_log_every_input(hm):
    """1 - u takes only the values k 2^-23, k = 1..2^23: check the Box-Muller log on all of them."""
    k = np.arange(1, (1 << 23) + 1, dtype=np.float64)
    x = (k * 2.0 ** -23).astype(np.float32)
    got = call1(hm, "f_bm_log", x)
    ref = np.log(x.astype(np.float64))
    assert got[-1] == 0.0
    assert ulps(got[:-1], ref[:-1]).max() < 0.95
    assert np.all(got <= 0)


def test_box_muller_pair(hm):
    rng = np.random.default_rng(8)
    n = 1_000_000
    u0 = ((rng.integers(0, 1 << 23, n)).astype(np.float64) * 2.0 ** -23).astype(np.float32)
    u1 = ((rng.integers(0, 1 << 23, n)).astype(np.float64) * 2.0 ** -23).astype(np.float32)
    u0[:3] = [0.0, 1 - 2.0 ** -23, 2.0 ** -23]
    o = np.empty(2 * n, np.float32)
    hm.f_bm_pair(C.c_void_p(u0.ctypes.data), C.c_void_p(u1.ctypes.data), C.c_void_p(o.ctypes.data), C.c_long(n),
                 C.c_float(0.5), C.c_float(2.0))
    r = np.sqrt(-2 * np.log(1 - u0.astype(np.float64))) * 2.0
    ang = (np.float32(6.28318530718) * u1).astype(np.float64)
    assert o[0] == 0.5 and o[1] == 0.5                                   # u0 = 0: r = 0 exactly
    assert np.abs(o[0::2] - (0.5 + r * np.sin(ang))).max() < 2e-6
    assert np.abs(o[1::2] - (0.5 + r * np.cos(ang))).max() < 2e-6


def test_asinh(hm):
    rng = np.random.default_rng(9)
    x = np.concatenate([rng.uniform(-4, 4, 1_000_000), np.exp(rng.uniform(-90, 88, 1_000_000)) * rng.choice([-1, 1], 1_000_000),
                        [0.0, -0.0, 1e-45, -1e-40, 1e18, 9.9e17, -3e38, 0.5, 1.0]]).astype(np.float32)
    got = call1(hm, "f_asinh", x)
    ref = np.arcsinh(x.astype(np.float64))
    nz = x != 0
    assert ulps(got[nz], ref[nz]).max() < 2.5
    assert np.all(np.signbit(got) == np.signbit(x))
    sp = call1(hm, "f_asinh", [np.inf, -np.inf, np.nan])
    assert sp[0] == np.inf and sp[1] == -np.inf and np.isnan(sp[2])


@pytest.mark.parametrize("s", [2.7, 0.5, -1.5, 3.0, 7.7, -7.75, 1 / 3, 1e-3, -0.25, 5.5, 1.0, 0.0, 6.3])
def test_pow_scalar_binomial(hm, s):
    """pows_core (the a ** s kernel for a launch-constant s): 2^(s e) * rc^-s * binomial series, <= 0.5001 ulp
    over the whole positive normal range; overflow -> inf, underflow -> 0 come out of the final conversion."""
    rs = np.random.default_rng(11)
    n = 400_000
    x = np.concatenate([rs.uniform(0.5, 2, n), np.exp(rs.uniform(-87, 88, n)), 1 + rs.uniform(-1e-2, 1e-2, n),
                        rs.uniform(0.9, 1.4, n)]).astype(np.float32)
    s32 = np.float32(s)
    hm.p_plan.argtypes = [C.c_float]
    assert hm.p_plan(s32) in (6, 8, 10)
    z = np.empty_like(x)
    hm.p_pows(C.c_void_p(x.ctypes.data), C.c_float(s32), C.c_void_p(z.ctypes.data), C.c_long(x.size))
    with np.errstate(all="ignore"):
        ex = np.power(x.astype(np.float64), np.float64(s32))
    ok = np.isfinite(ex) & (np.abs(ex) > 1.2e-38) & (np.abs(ex) < 3.4e38)
    assert ulps(z[ok], ex[ok]).max() <= 0.5001
    assert np.all(np.isinf(z[ex > 3.5e38])) and np.all(z[ex < 1.1e-38] == 0)
    # the reference's own points and exact cases
    if s == 2.7:
        z3 = np.empty(3, np.float32)
        x3 = np.float32([1, 2, 3])
        hm.p_pows(C.c_void_p(x3.ctypes.data), C.c_float(s32), C.c_void_p(z3.ctypes.data), C.c_long(3))
        assert bits(z3) == ["0x3f800000", "0x40cfefc6", "0x419b5a2a"]


def test_pow_scalar_binomial_plan(hm):
    hm.p_plan.argtypes = [C.c_float]
    assert hm.p_plan(2.7) == 6 and hm.p_plan(7.7) == 8
    assert hm.p_plan(8.0) == 0 and hm.p_plan(float("nan")) == 0 and hm.p_plan(float("inf")) == 0
