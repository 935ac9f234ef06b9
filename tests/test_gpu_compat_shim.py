"""The INTEGRATION.md shim as shipped code: `vulkpy_b200.compat._vkarray` has the surface of the
reference's pybind11 module (_vkarray.cc:756-898) and drives the CUDA kernels through the C ABI alone.
Every call below is one the reference's Python layer makes (vkarray.py:26-44,110-119,421-431;
random.py:230-312), checked against the oracle."""
import numpy as np
import pytest

from oracle import vulkpy_oracle as orc

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def shim():
    from vulkpy_b200.compat import _vkarray
    return _vkarray


def test_buffers_submit_job_wait(shim):
    gpu = shim.createGPU(0, 0.0)
    assert gpu.canSubgroupArithmetic() is True
    rs = np.random.default_rng(0)
    a_h, b_h = rs.uniform(0.5, 2, 1000).astype(F), rs.uniform(-2, 2, 1000).astype(F)
    a, b = gpu.toBuffer(a_h), gpu.toBuffer(b_h)
    c = gpu.createBuffer(1000)
    assert c.size() == 1000 and c.info() is c and c.range() is c
    # the reference names a kernel by the path of its .spv (util.py:58-72)
    job = gpu.submit("/site-packages/vulkpy/shader/add.spv", 64, 1, 1, [a.info(), b.info(), c.info()],
                     shim.DataShape(1000, 1, 1), shim.VectorParams(1000), [])
    job.wait()
    np.testing.assert_array_equal(np.asarray(c), orc.binary("add", a_h, b_h))
    job = gpu.submit("mul_scalar.spv", 64, 1, 1, [a, c], shim.DataShape(1000, 1, 1), shim.VectorScalarParams(1000, 2.5), [job])
    job.wait(10_000_000_000)
    np.testing.assert_array_equal(np.asarray(c), orc.scalar("mul", a_h, 2.5))
    gpu.submit("iexp", 64, 1, 1, [c], shim.DataShape(1000, 1, 1), shim.VectorParams(1000), []).wait()
    np.testing.assert_allclose(np.asarray(c), np.exp((a_h * F(2.5)).astype(np.float64)), rtol=1.2e-7)
    gpu.flush([c.range()])
    gpu.wait()
    with pytest.raises(RuntimeError, match="Unknown Operation"):
        gpu.submit("no_such_shader.spv", 64, 1, 1, [a], shim.DataShape(1, 1, 1), shim.VectorParams(1), [])
    # host writes through the NumPy view are seen by the next kernel (coherent mapping of the reference)
    view = np.asarray(a)
    view[:] = 3.0
    gpu.submit("add", 64, 1, 1, [a, b, c], shim.DataShape(1000, 1, 1), shim.VectorParams(1000), []).wait()
    np.testing.assert_array_equal(np.asarray(c), F(3.0) + b_h)


def test_reductions_matmul_broadcast_through_the_shim(shim):
    gpu = shim.createGPU(0, 0.0)
    rs = np.random.default_rng(1)
    x_h = rs.uniform(0, 1, (6, 5, 4)).astype(F)
    x = gpu.toBuffer(x_h)
    out = gpu.createBuffer(24)
    gpu.submit("sum_axis.spv", 1, 64, 1, [x, out], shim.DataShape(6, 4, 1), shim.AxisReductionParams(6, 5, 4), []).wait()
    np.testing.assert_allclose(np.asarray(out).reshape(6, 4), x_h.astype(np.float64).sum(axis=1), rtol=2e-6)
    m_h, n_h = rs.uniform(-1, 1, (7, 9)).astype(F), rs.uniform(-1, 1, (9, 3)).astype(F)
    c = gpu.createBuffer(21)
    gpu.submit("matmul.spv", 1, 64, 1, [gpu.toBuffer(m_h), gpu.toBuffer(n_h), c], shim.DataShape(7, 3, 1),
               shim.MatMulParams(7, 9, 3), []).wait()
    np.testing.assert_allclose(np.asarray(c).reshape(7, 3), orc.matmul(m_h, n_h), atol=1e-5)
    # add_broadcast: bindings A, B, C, shapes (uint32 [shapeA | shapeB | shapeC], vkarray.py:492-519)
    a_h, b_h = rs.uniform(0, 1, (1, 2, 2)).astype(F), rs.uniform(0, 1, (2, 2, 1)).astype(F)
    sh = gpu.toU32Buffer(np.array([1, 2, 2, 2, 2, 1, 2, 2, 2], np.uint32))
    o = gpu.createBuffer(8)
    gpu.submit("add_broadcast.spv", 64, 1, 1, [gpu.toBuffer(a_h), gpu.toBuffer(b_h), o, sh], shim.DataShape(8, 1, 1),
               shim.Multi3BroadcastParams(4, 4, 8, 3), []).wait()
    np.testing.assert_array_equal(np.asarray(o).reshape(2, 2, 2), a_h + b_h)
    idx = gpu.toU32Buffer(np.array([5, 0, 119], np.uint32))
    g = gpu.createBuffer(3)
    gpu.submit("gather.spv", 64, 1, 1, [x, idx, g], shim.DataShape(3, 1, 1), shim.VectorParams(3), []).wait()
    np.testing.assert_array_equal(np.asarray(g), x_h.reshape(-1)[[5, 0, 119]])


def test_xoshiro_through_the_shim(shim):
    gpu = shim.createGPU(0, 0.0)
    rng = shim.Xoshiro128pp(gpu, "prng_xoshiro128pp_uint32.spv", "prng_xoshiro128pp_float.spv", 64, 0)
    o = orc.Xoshiro128pp(64, 0)
    f = gpu.createBuffer(3)
    rng.random_float(3, f.info()).wait()
    np.testing.assert_array_equal(np.asarray(f), o.random(3))
    np.testing.assert_allclose(np.asarray(f), [0.42977667, 0.8235899, 0.90622926], rtol=1e-7)   # random.py:12-24
    u = gpu.createU32Buffer(1000)
    rng.random_uint32(1000, u.info()).wait()
    np.testing.assert_array_equal(np.asarray(u), o.randint(1000))
    unseeded = shim.Xoshiro128pp(gpu, "", "", 8)
    unseeded.random_float(3, f).wait()
    assert ((0 <= np.asarray(f)) & (np.asarray(f) < 1)).all()
