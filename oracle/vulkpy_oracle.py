"""
CPU oracle for the vulkpy array hot path -- TEST INFRASTRUCTURE ONLY.

A NumPy restatement of what the reference's shaders and PRNG compute, written from the
reference sources (each function cites the file:line it follows).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it; the product (``vulkpy_b200``) never does.

Pinning: the reference cannot be built or imported in this image (it needs
libvulkan, glslc and a Vulkan ICD), so the oracle is pinned against the value-level vectors
the reference itself publishes -- the ``Xoshiro128pp(seed=0)`` docstring of
vulkpy/random.py:12-24 and the known answers of test/test_vulkpy.py, test/test_nn.py and
doc/broadcasting.md -- see tests/test_oracle.py and tests/golden/ -- and, wholesale, against the
reference's own 233 unit tests, which tests/test_reference_suite_cpu.py runs unchanged from
/root/reference/test with this oracle answering every kernel behind the Python layer.
GLSL built-ins (exp, log, pow, sin, ...) are implemented by the Vulkan driver, which the
reference does not pin (Dockerfile:1-10); for those the oracle is the correctly rounded
float32 value of the float64 result, and tests state their tolerance: "parity unpinned
beyond the reference's own 3-point tests" for transcendentals.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
U32 = np.uint32
MASK32 = 0xFFFFFFFF
MASK64 = 0xFFFFFFFFFFFFFFFF


# --------------------------------------------------------------------------------------------
# PRNG: vulkpy/_vkarray.cc:577-719, shader/prng_xoshiro128pp_uint32.comp, ..._float.comp
# --------------------------------------------------------------------------------------------
def splitmix64(x: int) -> int:
    """_vkarray.cc:589-595"""
    x = (x + 0x9e3779b97f4a7c15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xbf58476d1ce4e5b9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94d049bb133111eb) & MASK64
    return z ^ (z >> 31)


def _rotl(x: int, k: int) -> int:
    return ((x << k) | (x >> (32 - k))) & MASK32


def next_scalar(s: list) -> int:
    """One xoshiro128++ step on a 4-word python list, in place (_vkarray.cc:627-641)."""
    result = (_rotl((s[0] + s[3]) & MASK32, 7) + s[0]) & MASK32
    t = (s[1] << 9) & MASK32
    s[2] ^= s[0]
    s[3] ^= s[1]
    s[1] ^= s[2]
    s[0] ^= s[3]
    s[2] ^= t
    s[3] = _rotl(s[3], 11)
    return result


JUMP = (0x8764000b, 0xf542d2d3, 0x6fa035c3, 0x77f2db5b)


def jump_reference(s: list) -> None:
    """The reference's jump (_vkarray.cc:597-621): the accumulators s0..s3 are written back to
    ``s`` after EACH of the four JUMP words and are never reset (unlike Vigna's jump())."""
    acc = [0, 0, 0, 0]
    for j in JUMP:
        for b in range(32):
            if j & (1 << b):
                for i in range(4):
                    acc[i] ^= s[i]
            next_scalar(s)
        s[:] = acc


def seed_states(size: int, seed: int) -> np.ndarray:
    """Initial state buffer [size, 4] (_vkarray.cc:653-674): four chained splitmix64 outputs
    truncated to 32 bits form lane 0, lane i = jump(lane i-1)."""
    s = []
    for _ in range(4):
        seed = splitmix64(seed)
        s.append(seed & MASK32)
    out = np.empty((size, 4), dtype=U32)
    out[0] = s
    for i in range(1, size):
        jump_reference(s)
        out[i] = s
    return out


def next_lanes(state: np.ndarray, nlanes: int) -> np.ndarray:
    """Advance lanes [0, nlanes) of ``state`` ([size,4] uint32) once and return their draws
    (shader/prng_xoshiro128pp_uint32.comp:26-43)."""
    s = state[:nlanes]
    s0, s1, s2, s3 = (s[:, i].copy() for i in range(4))
    tmp = s0 + s3
    result = ((tmp << U32(7)) | (tmp >> U32(25))) + s0
    t = s1 << U32(9)
    s2 ^= s0
    s3 ^= s1
    s1 ^= s2
    s0 ^= s3
    s2 ^= t
    s3 = (s3 << U32(11)) | (s3 >> U32(21))
    state[:nlanes, 0], state[:nlanes, 1], state[:nlanes, 2], state[:nlanes, 3] = s0, s1, s2, s3
    return result


def u32_to_unit_float(r: np.ndarray) -> np.ndarray:
    """uintBitsToFloat((r >> 9) | 0x3f800000) - 1.0 (shader/prng_xoshiro128pp_float.comp:33)."""
    bits = ((r >> U32(9)) | U32(0x3f800000)).astype(U32)
    return (bits.view(F32) - F32(1.0)).astype(F32)


class Xoshiro128pp:
    """Stream semantics of PRNG::Xoshiro128pp::random (_vkarray.cc:697-717): chunks of ``size``,
    chunk c writes out[c*size + lane], the last chunk advances the first n % size lanes only,
    state persists between calls."""

    def __init__(self, size: int = 64, seed: int = 0):
        self.size = size
        self.state = seed_states(size, seed)

    def randint(self, n: int) -> np.ndarray:
        out = np.empty(n, dtype=U32)
        with np.errstate(over="ignore"):
            if n <= self.size:
                out[:] = next_lanes(self.state, n)
                return out
            for i in range(0, n, self.size):
                m = min(self.size, n - i)
                out[i:i + m] = next_lanes(self.state, m)
        return out

    def random(self, n: int) -> np.ndarray:
        return u32_to_unit_float(self.randint(n))

    def normal(self, n: int, mean: float = 0.0, stddev: float = 1.0) -> np.ndarray:
        """vulkpy/random.py:105-124: n uniforms in place for even n, n+1 for odd n."""
        u = self.random(n if n % 2 == 0 else n + 1)
        return box_muller(u, n, mean, stddev)

    def randrange(self, n: int, low: int, high: int) -> np.ndarray:
        """vulkpy/random.py:157-186 (high is exclusive; the shader gets high-1)."""
        if low == 0 and high == (1 << 32):
            return self.randint(n)
        return randrange_shader(self.random(n), low, high - 1)


def box_muller(u: np.ndarray, n: int, mean: float, stddev: float) -> np.ndarray:
    """shader/prng_box_muller.comp:19-32 with one float32 rounding per operation; log, sqrt,
    sin, cos are taken correctly rounded (driver built-ins in the reference)."""
    mean, stddev = F32(mean), F32(stddev)
    u = u.astype(F32)
    a, b = u[0::2], u[1::2]
    one_minus = (F32(1.0) - a).astype(F32)
    lg = np.log(one_minus.astype(np.float64)).astype(F32)
    with np.errstate(invalid="ignore"):
        r = (np.sqrt((F32(-2.0) * lg).astype(F32).astype(np.float64)).astype(F32) * stddev).astype(F32)
    angle = (F32(6.28318530718) * b).astype(F32)
    s = np.sin(angle.astype(np.float64)).astype(F32)
    c = np.cos(angle.astype(np.float64)).astype(F32)
    out = np.empty(2 * len(a), dtype=F32)
    out[0::2] = (mean + (r * s).astype(F32)).astype(F32)
    out[1::2] = (mean + (r * c).astype(F32)).astype(F32)
    return out[:n]


def randrange_shader(u: np.ndarray, low: int, high_inclusive: int) -> np.ndarray:
    """b = low + uint(float(high - low + 1) * a) (shader/prng_randrange.comp:20-27)."""
    rng = F32((high_inclusive - low + 1) & MASK32)
    prod = (rng * u.astype(F32)).astype(F32)
    return (U32(low) + prod.astype(np.uint64).astype(U32)).astype(U32)


# --------------------------------------------------------------------------------------------
# element-wise families: shader/add.comp:21-26 etc.
# --------------------------------------------------------------------------------------------
def _cr(f, *xs):
    """Correctly rounded float32 of a float64 evaluation."""
    with np.errstate(all="ignore"):
        return np.asarray(f(*[np.asarray(x, dtype=F32).astype(np.float64) for x in xs])).astype(F32)


def _ftz(r):
    """Flush float32 subnormal results to zero: GPU Vulkan drivers do, and the reference's own
    test relies on it (test/test_nn.py:120-126 expects softmax([100, 0]) == [1, 0] exactly, i.e.
    exp(-100) = 3.8e-44 must be 0)."""
    r = np.asarray(r, dtype=F32)
    return np.where(np.abs(r) < F32(1.17549435e-38), np.copysign(F32(0), r), r).astype(F32)


def _glsl_sign(x):
    return np.where(x > 0, 1.0, np.where(x < 0, -1.0, 0.0))


BINARY = {
    "add": lambda a, b: (a + b).astype(F32),
    "sub": lambda a, b: (a - b).astype(F32),
    "mul": lambda a, b: (a * b).astype(F32),
    "div": lambda a, b: (a / b).astype(F32),
    "max": lambda a, b: np.maximum(a, b).astype(F32),
    "min": lambda a, b: np.minimum(a, b).astype(F32),
    "pow": lambda a, b: _ftz(_cr(np.power, a, b)),
}

UNARY = {
    "abs": lambda a: np.abs(a).astype(F32),
    "sign": lambda a: _glsl_sign(a).astype(F32),
    "sin": lambda a: _cr(np.sin, a), "cos": lambda a: _cr(np.cos, a), "tan": lambda a: _cr(np.tan, a),
    "asin": lambda a: _cr(np.arcsin, a), "acos": lambda a: _cr(np.arccos, a),
    "atan": lambda a: _cr(np.arctan, a),
    "sinh": lambda a: _cr(np.sinh, a), "cosh": lambda a: _cr(np.cosh, a), "tanh": lambda a: _cr(np.tanh, a),
    "asinh": lambda a: _cr(np.arcsinh, a), "acosh": lambda a: _cr(np.arccosh, a),
    "atanh": lambda a: _cr(np.arctanh, a),
    "exp": lambda a: _ftz(_cr(np.exp, a)), "log": lambda a: _cr(np.log, a),
    "exp2": lambda a: _ftz(_cr(np.exp2, a)), "log2": lambda a: _cr(np.log2, a),
    "sqrt": lambda a: _cr(np.sqrt, a),
    "invsqrt": lambda a: _cr(lambda x: 1.0 / np.sqrt(x), a),
}


def binary(op: str, a, b):
    """c[i] = a[i] op b[i] in float32 (shader/add.comp:21-26 and siblings)."""
    with np.errstate(all="ignore"):
        return BINARY[op](np.asarray(a, dtype=F32), np.asarray(b, dtype=F32))


def scalar(op: str, a, s, reverse: bool = False):
    """b[i] = a[i] op s, or s op a[i] for the r-forms; the scalar crosses the boundary as a C
    float (shader/add_scalar.comp:19-24, rsub_scalar.comp:23; _vkarray.cc:841-845)."""
    a = np.asarray(a, dtype=F32)
    sv = np.full_like(a, F32(s))
    return binary(op, sv, a) if reverse else binary(op, a, sv)


def unary(op: str, a):
    """b[i] = f(a[i]) (shader/abs.comp:22 ... shader/invsqrt.comp:22)."""
    return UNARY[op](np.asarray(a, dtype=F32))


def clamp(a, lo, hi):
    """GLSL clamp(x, lo, hi) = min(max(x, lo), hi) (shader/clamp.comp:24-29)."""
    a = np.asarray(a, dtype=F32)
    return np.minimum(np.maximum(a, np.asarray(lo, dtype=F32)), np.asarray(hi, dtype=F32)).astype(F32)


def cross_entropy(x, y):
    """L = -y * log(x + 1e-8) (shader/nn_cross_entropy.comp:25)."""
    x, y = np.asarray(x, dtype=F32), np.asarray(y, dtype=F32)
    return ((-y) * _cr(np.log, (x + F32(1e-8)).astype(F32))).astype(F32)


def cross_entropy_backward(x, y):
    """dx = -y / (x + 1e-8) (shader/nn_cross_entropy_backward.comp:25)."""
    x, y = np.asarray(x, dtype=F32), np.asarray(y, dtype=F32)
    with np.errstate(all="ignore"):
        return ((-y) / (x + F32(1e-8)).astype(F32)).astype(F32)


# --------------------------------------------------------------------------------------------
# broadcasting: shader/add_broadcast.comp:25-45, iadd_broadcast.comp:22-41, broadcast.comp:25-44
# --------------------------------------------------------------------------------------------
def broadcast_indices(shapes, out_shape):
    """Literal restatement of the shader's index loop, vectorised over the output index ``ci``:
    for every operand returns the flat source index of each output element."""
    ndim = len(out_shape)
    padded = [(1,) * (ndim - len(s)) + tuple(s) for s in shapes]
    nout = int(np.prod(out_shape, dtype=np.int64))
    ci = np.arange(nout, dtype=np.int64)
    sizes = [int(np.prod(s, dtype=np.int64)) for s in padded]
    size_c = nout
    idx = [np.zeros(nout, dtype=np.int64) for _ in shapes]
    rem = ci.copy()
    for dim in range(ndim):
        for k in range(len(shapes)):
            sizes[k] //= padded[k][dim]
        size_c //= out_shape[dim]
        d = rem // size_c
        for k in range(len(shapes)):
            idx[k] += sizes[k] * np.minimum(d, padded[k][dim] - 1)
        rem = rem % size_c
    return idx


def broadcast_binary(op: str, a, b):
    a, b = np.asarray(a, dtype=F32), np.asarray(b, dtype=F32)
    out_shape = np.broadcast_shapes(a.shape, b.shape)
    ia, ib = broadcast_indices([a.shape, b.shape], out_shape)
    return binary(op, a.reshape(-1)[ia], b.reshape(-1)[ib]).reshape(out_shape)


def broadcast_to(a, shape):
    a = np.asarray(a, dtype=F32)
    (ia,) = broadcast_indices([a.shape], tuple(shape))
    return a.reshape(-1)[ia].reshape(shape)


# --------------------------------------------------------------------------------------------
# reductions: shader/sum.comp:18-30, sum_axis.comp:20-32, sum_axis_rebroadcast.comp:20-35
# --------------------------------------------------------------------------------------------
_RED = {
    "sum": (lambda acc, x: (acc + x).astype(F32), lambda first: np.zeros_like(first)),
    "prod": (lambda acc, x: (acc * x).astype(F32), lambda first: np.ones_like(first)),
    "maximum": (lambda acc, x: np.maximum(acc, x).astype(F32), lambda first: first.copy()),
    "minimum": (lambda acc, x: np.minimum(acc, x).astype(F32), lambda first: first.copy()),
}


def reduce_axis(op: str, a, axis: int, rebroadcast: bool = False):
    """One thread per output, serial loop k = 0..axis_size-1 in float32 (shader/sum_axis.comp:27-31;
    maximum/minimum seed with the first element, maximum_axis.comp)."""
    a = np.asarray(a, dtype=F32)
    axis = axis % a.ndim
    step, init = _RED[op]
    moved = np.moveaxis(a, axis, 0)
    acc = init(moved[0])
    for k in range(moved.shape[0]):
        acc = step(acc, moved[k])
    if rebroadcast:
        return np.broadcast_to(np.expand_dims(acc, axis), a.shape).copy()
    return acc


def reduce_full_reference(op: str, a):
    """The fallback full reduction exactly as dispatched (vkarray.py:1246-1274 + shader/sum.comp):
    passes of one 64-thread workgroup, thread i folding a[i::m].  Only defined for n <= 4096."""
    v = np.asarray(a, dtype=F32).reshape(-1)
    assert v.size <= 4096, "reference result is undefined beyond 4096 elements (SURVEY Q2)"
    step, init = _RED[op]
    while True:
        m = (v.size + 63) // 64
        out = np.empty(m, dtype=F32)
        for i in range(m):
            col = v[i::m]
            acc = init(col[:1])[0]
            for x in col:
                acc = step(np.asarray(acc), np.asarray(x))
            out[i] = acc
        if m == 1:
            return out
        v = out


def reduce_full_exact(op: str, a):
    """Mathematical definition in float64 (the yardstick where the reference is undefined)."""
    a = np.asarray(a, dtype=np.float64)
    return {"sum": np.sum, "prod": np.prod, "maximum": np.max, "minimum": np.min}[op](a)


# --------------------------------------------------------------------------------------------
# gather: shader/gather.comp:21-26, gather_axis.comp:24-43
# --------------------------------------------------------------------------------------------
def gather(a, idx):
    return np.asarray(a, dtype=F32).reshape(-1)[np.asarray(idx, dtype=np.int64)]


def gather_axis(a, idx, axis: int):
    """c[k, i, j] = a[i, b[k], j]: index dimensions lead (vkarray.py:1506-1517)."""
    a = np.asarray(a, dtype=F32)
    idx = np.asarray(idx, dtype=np.int64)
    prev, post = a.shape[:axis], a.shape[axis + 1:]
    flat = a.reshape(int(np.prod(prev, dtype=np.int64)), a.shape[axis], int(np.prod(post, dtype=np.int64)))
    out = np.stack([flat[:, k, :] for k in idx.reshape(-1)], axis=0)
    return out.reshape(idx.shape + prev + post)


# --------------------------------------------------------------------------------------------
# contractions: shader/matmul.comp:23-33, batch_affine.comp:25-39
# --------------------------------------------------------------------------------------------
def matmul(a, b):
    """Serial float32 accumulation over k, product rounded before the add (no FMA assumed)."""
    a, b = np.asarray(a, dtype=F32), np.asarray(b, dtype=F32)
    a2 = a.reshape(1, -1) if a.ndim == 1 else a
    b2 = b.reshape(-1, 1) if b.ndim == 1 else b
    acc = np.zeros((a2.shape[0], b2.shape[1]), dtype=F32)
    for s in range(a2.shape[1]):
        acc = (acc + (a2[:, s:s + 1] * b2[s:s + 1, :]).astype(F32)).astype(F32)
    shape = a.shape[:-1] + b.shape[1:]
    return acc.reshape(shape if shape else (1,))


def batch_affine(w, bias, x):
    """y[b,o] = sum_i w[o,i] x[b,i] + bias[o] (shader/batch_affine.comp:33-38)."""
    w, bias, x = (np.asarray(v, dtype=F32) for v in (w, bias, x))
    acc = np.zeros((x.shape[0], w.shape[0]), dtype=F32)
    for i in range(x.shape[1]):
        acc = (acc + (x[:, i:i + 1] * w[:, i][None, :]).astype(F32)).astype(F32)
    return (acc + bias[None, :]).astype(F32)


# ---- argmax / argmin / permutation (additive; the reference does them with NumPy on the host:
# example/02-nn.py:82 `rng.shuffle(idx)`, :96 `np.argmax(pred_y, axis=1)`) -------------------------
def argmax(a, axis=None):
    a = np.asarray(a, dtype=np.float32)
    r = np.argmax(a, axis=axis)
    return np.asarray(r, dtype=np.uint32).reshape((1,) if axis is None or a.ndim == 1 else r.shape)


def argmin(a, axis=None):
    a = np.asarray(a, dtype=np.float32)
    r = np.argmin(a, axis=axis)
    return np.asarray(r, dtype=np.uint32).reshape((1,) if axis is None or a.ndim == 1 else r.shape)


def permutation_from_keys(keys):
    """Indices 0..n-1 stably sorted by uint32 keys (keys = the generator's next n randint draws)."""
    return np.argsort(np.asarray(keys, dtype=np.uint32), kind="stable").astype(np.uint32)
