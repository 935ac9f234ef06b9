"""Copy-engine transfers (Array.from_host / Array.to_host(wait=False)): ordering against the compute
stream, block recycling and the host view.  The reference has no asynchronous copies
(Buffer::set is a memcpy, _vkarray.cc:98-108); results must equal the synchronous path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vk():
    import vulkpy_b200 as vk
    return vk


@pytest.fixture(scope="module")
def gpu(vk):
    return vk.GPU(0)


def pinned(vk, data):
    h = vk.pinned_empty(np.shape(data), dtype=np.asarray(data).dtype)
    h[...] = data
    return h


def test_from_host_roundtrip(vk, gpu):
    x = np.random.default_rng(0).standard_normal((513, 257)).astype(np.float32)
    a = vk.Array.from_host(gpu, pinned(vk, x))
    assert a.shape == x.shape
    np.testing.assert_array_equal(np.asarray(a), x)                 # host view waits for the upload
    b = vk.Array.from_host(gpu, pinned(vk, x))
    out = vk.pinned_empty(x.shape)
    c = b * 2.0
    c.to_host(out, wait=False)
    c.wait()
    np.testing.assert_array_equal(out, x * np.float32(2))
    np.testing.assert_array_equal(b.to_host(), x)                   # synchronous download after async upload


def test_u32_from_host(vk, gpu):
    idx = np.arange(1000, dtype=np.uint32)[::-1].copy()
    u = vk.U32Array.from_host(gpu, pinned(vk, idx))
    t = vk.Array(gpu, data=np.arange(1000, dtype=np.float32))
    np.testing.assert_array_equal(np.asarray(t.gather(u)), idx.astype(np.float32))


def test_pageable_source_is_rejected(vk, gpu):
    with pytest.raises(RuntimeError, match="page-locked"):
        vk.Array.from_host(gpu, np.zeros(16, np.float32))
    a = vk.Array(gpu, data=np.zeros(16, np.float32))
    with pytest.raises(RuntimeError, match="page-locked"):
        a.to_host(np.zeros(16, np.float32), wait=False)
        a.wait()


def test_download_then_overwrite(vk, gpu):
    """Kernels bound to the array after an asynchronous download must not overtake it."""
    n = 1 << 24
    x = np.random.default_rng(1).random(n, dtype=np.float32)
    a = vk.Array(gpu, data=x)
    out = vk.pinned_empty((n,))
    a.to_host(out, wait=False)
    job = a.job
    a += 1.0
    a.wait()
    job.wait()
    np.testing.assert_array_equal(out, x)
    np.testing.assert_array_equal(a.to_host(), x + np.float32(1))


def test_upload_into_recycled_block(vk, gpu):
    """A freed block that enqueued kernels still read must not be overwritten early."""
    n = 1 << 24
    rng = np.random.default_rng(2)
    x, y = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    yh = pinned(vk, y)
    for _ in range(3):
        a = vk.Array(gpu, data=x)
        b = a * 3.0          # reads a
        c = b + a            # reads a again
        del a                # block goes back to the pool with work in flight
        d = vk.Array.from_host(gpu, yh)
        e = d + c
        np.testing.assert_array_equal(e.to_host(), y + (x * np.float32(3) + x))


def test_pipelined_steps(vk, gpu):
    """The loop bench.py's end-to-end leg runs: upload(i+1) overlaps compute(i) and download(i)."""
    n = 1 << 22
    rng = np.random.default_rng(3)
    steps = 6
    xs = [rng.random(n, dtype=np.float32) for _ in range(steps)]
    ins = [vk.pinned_empty((n,)) for _ in range(2)]
    outs = [vk.pinned_empty((n,)) for _ in range(2)]
    ins[0][...] = xs[0]
    nxt = vk.Array.from_host(gpu, ins[0])
    got = []
    for i in range(steps):
        cur = nxt
        r = (cur * 2.0 + 1.0).sqrt()
        r.to_host(outs[i % 2], wait=False)
        if i + 1 < steps:
            cur.wait()                      # upload i finished (r depends on it) -> ins[(i+1)%2] is free: it fed step i-1
            ins[(i + 1) % 2][...] = xs[i + 1]
            nxt = vk.Array.from_host(gpu, ins[(i + 1) % 2])
        r.wait()
        got.append(outs[i % 2].copy())
    for i in range(steps):
        np.testing.assert_array_equal(got[i], np.sqrt(xs[i] * np.float32(2) + np.float32(1)))
    gpu.wait()
