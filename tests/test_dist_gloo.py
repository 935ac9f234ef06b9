"""world_size-2 (and 3) gloo tests of the multi-GPU host logic in vulkpy_b200.dist: shard bounds,
which operations communicate and how much, global means, the sharded matmul / data-parallel
algorithms -- against the unsharded NumPy answer.  Local arrays and the transport are the CPU
stand-ins of tests/dist_sim.py; on the GPU box the same code runs over vk.Array + NCCL
(scripts/dist_check.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as td
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

F = np.float32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn_name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vulkpy_b200 import dist
        import dist_sim
        g = dist.Group(dist_sim.GlooTransport(rank, world), rank, world)
        globals()[fn_name](g, dist, dist_sim)
        if fn_name in ("body_reductions", "body_uneven"):     # again over the fused reduce + exchange seam
            g = dist.Group(dist_sim.GlooFusedTransport(rank, world), rank, world)
            globals()[fn_name](g, dist, dist_sim)
        q.put((rank, "ok"))
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        td.destroy_process_group()


def run(fn_name, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status in results:
        assert status == "ok", f"rank {rank}:\n{status}"


def _full(shape, seed=0, lo=0.5, hi=2.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(F)


def _shard(g, dist_sim, full):
    lo, hi = g.bounds(full.shape[0])
    from vulkpy_b200.dist import ShardedArray
    return ShardedArray(g, dist_sim.NumpyLocal(full[lo:hi]), full.shape)


# ---- bodies executed on every rank ---------------------------------------------------------------
def body_elementwise(g, dist, sim):
    a_f, b_f = _full((10, 6), 1), _full((10, 6), 2)
    row, col_f = _full((6,), 3), _full((10, 1), 4)
    a, b, col = _shard(g, sim, a_f), _shard(g, sim, b_f), _shard(g, sim, col_f)
    r = (a + b) * 2.0 - a / b
    r = r.max(b).clamp(0.5, 5.0)
    r = r + sim.NumpyLocal(row)          # replicated row vector
    r = r * col                          # (rows, 1) operand sharded like the rows
    r += b
    r = r.exp().log().sqrt()
    want = np.sqrt(np.log(np.exp((np.clip(np.maximum((a_f + b_f) * 2 - a_f / b_f, b_f), 0.5, 5) + row) * col_f + b_f)))
    np.testing.assert_allclose(r.to_numpy(), want, rtol=1e-5)
    # nothing above may communicate except the final to_numpy()
    assert [c[0] for c in g.t.calls] == ["allgather"]
    with pytest.raises(ValueError):
        a + sim.NumpyLocal(col_f)        # a full-leading-axis operand must be sharded


def body_reductions(g, dist, sim):
    a_f = _full((12, 5, 4), 5)
    a = _shard(g, sim, a_f)
    g.t.calls.clear()
    r = a.sum(axis=1)                                   # rows independent: no exchange
    assert g.t.calls == [] and r.shape == (12, 4)
    np.testing.assert_allclose(r.to_numpy(), a_f.sum(axis=1), rtol=1e-6)
    g.t.calls.clear()
    r = a.maximum(axis=(1, 2), keepdims=True)
    assert g.t.calls == [] and r.shape == (12, 1, 1)
    g.t.calls.clear()
    fused = hasattr(g.t, "reduce_allreduce")
    kind = "reduce_allreduce" if fused else "allreduce"
    r0 = a.sum(axis=0)                                  # one exchange of `post` floats
    assert g.t.calls == [(kind, 20)] and r0.shape == (5, 4)
    np.testing.assert_allclose(np.asarray(r0), a_f.sum(axis=0), rtol=1e-6)
    for name, f in [("sum", np.sum), ("prod", np.prod), ("maximum", np.max), ("minimum", np.min)]:
        g.t.calls.clear()
        r = getattr(a, name)()
        assert g.t.calls == [(kind, 1)] and r.shape == (1,)   # a single float crosses the wire
        np.testing.assert_allclose(np.asarray(r), [f(a_f.astype(np.float64))], rtol=1e-4)
    np.testing.assert_allclose(np.asarray(a.sum(axis=(0, 2))), a_f.sum(axis=(0, 2)), rtol=1e-6)
    assert a.sum(axis=(0, 2)).shape == (5,) and a.sum(axis=(0, 2), keepdims=True).shape == (1, 5, 1)
    assert a.maximum(axis=0, keepdims=True).shape == (1, 5, 4)
    np.testing.assert_allclose(np.asarray(a.mean()), [a_f.mean()], rtol=1e-6)            # global count
    np.testing.assert_allclose(np.asarray(a.mean(axis=0)), a_f.mean(axis=0), rtol=1e-6)
    np.testing.assert_allclose(a.mean(axis=2).to_numpy(), a_f.mean(axis=2), rtol=1e-6)
    rb = a.sum(axis=0, rebroadcast=True)
    np.testing.assert_allclose(rb.to_numpy(), np.broadcast_to(a_f.sum(axis=0, keepdims=True), a_f.shape), rtol=1e-6)
    rb = a.maximum(axis=2, rebroadcast=True)
    np.testing.assert_allclose(rb.to_numpy(), np.broadcast_to(a_f.max(axis=2, keepdims=True), a_f.shape))
    assert a.sum(keepdims=True).shape == (1, 1, 1)


def body_uneven(g, dist, sim):
    # 7 rows over 3 ranks: 3 + 2 + 2
    assert [dist.shard_bounds(7, 3, r) for r in range(3)] == [(0, 3), (3, 5), (5, 7)]
    a_f = _full((7, 3), 6)
    a = _shard(g, sim, a_f)
    np.testing.assert_allclose(np.asarray(a.sum()), [a_f.sum()], rtol=1e-6)
    np.testing.assert_allclose(np.asarray(a.mean(axis=0)), a_f.mean(axis=0), rtol=1e-6)
    with pytest.raises(ValueError):
        a.allgather()                                   # NCCL all-gather needs equal shards
    with pytest.raises(ValueError):
        dist.ShardedArray(g, sim.NumpyLocal(a_f), a_f.shape)   # wrong local block


def body_matmul(g, dist, sim):
    A_f, B_f = _full((8, 6), 7, -1, 1), _full((6, 4), 8, -1, 1)
    A, B = _shard(g, sim, A_f), _shard(g, sim, B_f)
    g.t.calls.clear()
    C = A @ B                                           # B sharded by K: all-gather, then local GEMM
    assert g.t.calls == [("allgather", B_f.size // g.world)]
    np.testing.assert_allclose(C.to_numpy(), A_f @ B_f, rtol=1e-5)
    g.t.calls.clear()
    C2 = A @ sim.NumpyLocal(B_f)                        # B replicated: no exchange
    assert g.t.calls == []
    np.testing.assert_allclose(C2.to_numpy(), A_f @ B_f, rtol=1e-5)
    idx = np.random.default_rng(9).integers(0, 48, 10)
    lo, hi = g.bounds(10)
    out = dist.gather_replicated(g, sim.NumpyLocal(A_f), idx[lo:hi], 10)
    np.testing.assert_allclose(out.to_numpy(), A_f.reshape(-1)[idx])


def body_matmul_pull(g, dist, sim):
    """A transport with `matmul_allgather` takes the sharded x sharded product (no all-gather call);
    when it declines the shape, ShardedArray falls back to all-gather + local GEMM."""
    g.t.__class__ = sim.GlooPullTransport
    A_f, B_f = _full((8 * g.world, 6 * g.world), 7, -1, 1), _full((6 * g.world, 4), 8, -1, 1)
    A, B = _shard(g, sim, A_f), _shard(g, sim, B_f)
    g.t.calls.clear()
    C = A @ B
    assert g.t.calls == [("pull", (B_f.size // g.world) * (g.world - 1))], g.t.calls
    assert C.shape == (A_f.shape[0], 4)
    np.testing.assert_allclose(C.to_numpy(), A_f @ B_f, rtol=1e-5, atol=1e-6)
    # declined (too few local rows): the all-gather path
    A2_f = _full((2 * g.world, 6 * g.world), 9, -1, 1)
    g.t.calls.clear()
    C2 = _shard(g, sim, A2_f) @ B
    assert [c[0] for c in g.t.calls] == ["allgather"], g.t.calls
    np.testing.assert_allclose(C2.to_numpy(), A2_f @ B_f, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        A @ _shard(g, sim, _full((4 * g.world, 4), 10))      # inner dimensions differ


def body_data_parallel(g, dist, sim):
    """Gradient all-reduce + 1/world scaling reproduces the single-process full-batch SGD step."""
    class P:  # minimal Parameter / layer / loss / Sequence surface used by DataParallel
        def __init__(self, v):
            self.value, self.grad = sim.NumpyLocal(v), sim.NumpyLocal(np.zeros_like(v))

    class Lin:
        def __init__(self, w):
            self.w = P(w)

    class Net:
        def __init__(self, w, lr):
            self.L, self.lr = (Lin(w),), lr

        def _forward(self, x):
            self._x = x
            return x @ sim.NumpyLocal(self.L[0].w.value.a.T)

        def loss(self, pred, y):
            self._d = pred - y
            return ((self._d * self._d).sum(axis=1)).sum(axis=0) * (1.0 / pred.shape[0])   # mean over the LOCAL batch

        def _zero_grad(self):
            self.L[0].w.grad = sim.NumpyLocal(np.zeros_like(self.L[0].w.value.a))

        def _backward(self):
            d = self._d.a * (2.0 / self._d.a.shape[0])
            self.L[0].w.grad += sim.NumpyLocal(d.T @ self._x.a)

        def _update(self):
            self.L[0].w.value += self.L[0].w.grad * (-self.lr)

    W = _full((3, 5), 10, -1, 1)
    X, Y = _full((8, 5), 11, -1, 1), _full((8, 3), 12, -1, 1)
    lo, hi = g.bounds(8)
    net = Net(W.copy(), 0.1)
    dp = dist.DataParallel(net, g)
    g.t.calls.clear()
    _, loss = dp.train(sim.NumpyLocal(X[lo:hi]), sim.NumpyLocal(Y[lo:hi]))
    assert [c for c in g.t.calls] == [("allreduce", 15), ("allreduce", 1)]   # one gradient tensor + the loss
    D = X @ W.T - Y
    want_W = W - 0.1 * (2.0 / 8) * D.T @ X
    np.testing.assert_allclose(net.L[0].w.value.a, want_W, rtol=1e-5)
    np.testing.assert_allclose(np.asarray(loss).reshape(-1)[0], (D * D).sum(axis=1).mean(), rtol=1e-5)


def body_data_parallel_bucket(g, dist, sim):
    """With a transport that offers `allreduce_many`, DataParallel sends gradients AND the loss as one
    bucket (one collective per step) and gets the same step as the per-tensor path."""
    g.t.__class__ = sim.GlooBucketTransport
    body_data_parallel_checked(g, dist, sim, [("allreduce_many", 16, 2)])


def body_data_parallel_checked(g, dist, sim, want_calls):
    class P:
        def __init__(self, v):
            self.value, self.grad = sim.NumpyLocal(v), sim.NumpyLocal(np.zeros_like(v))

    class Lin:
        def __init__(self, w):
            self.w = P(w)

    class Net:
        def __init__(self, w, lr):
            self.L, self.lr = (Lin(w),), lr

        def _forward(self, x):
            self._x = x
            return x @ sim.NumpyLocal(self.L[0].w.value.a.T)

        def loss(self, pred, y):
            self._d = pred - y
            return ((self._d * self._d).sum(axis=1)).sum(axis=0) * (1.0 / pred.shape[0])

        def _zero_grad(self):
            self.L[0].w.grad = sim.NumpyLocal(np.zeros_like(self.L[0].w.value.a))

        def _backward(self):
            d = self._d.a * (2.0 / self._d.a.shape[0])
            self.L[0].w.grad += sim.NumpyLocal(d.T @ self._x.a)

        def _update(self):
            self.L[0].w.value += self.L[0].w.grad * (-self.lr)

    W = _full((3, 5), 10, -1, 1)
    X, Y = _full((8, 5), 11, -1, 1), _full((8, 3), 12, -1, 1)
    lo, hi = g.bounds(8)
    net = Net(W.copy(), 0.1)
    dp = dist.DataParallel(net, g)
    g.t.calls.clear()
    _, loss = dp.train(sim.NumpyLocal(X[lo:hi]), sim.NumpyLocal(Y[lo:hi]))
    assert list(g.t.calls) == want_calls, g.t.calls
    D = X @ W.T - Y
    want_W = W - 0.1 * (2.0 / 8) * D.T @ X
    np.testing.assert_allclose(net.L[0].w.value.a, want_W, rtol=1e-5)
    np.testing.assert_allclose(np.asarray(loss).reshape(-1)[0], (D * D).sum(axis=1).mean(), rtol=1e-5)


# ---- pytest entry points ----------------------------------------------------------------------------
def test_shard_bounds():
    from vulkpy_b200.dist import shard_bounds
    for n in (0, 1, 7, 8, 16384):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


def test_elementwise_world2():
    run("body_elementwise", 2)


def test_reductions_world2():
    run("body_reductions", 2)


def test_reductions_world3():
    run("body_reductions", 3)


def test_uneven_world3():
    run("body_uneven", 3)


def test_matmul_and_gather_world2():
    run("body_matmul", 2)


def test_matmul_pull_world2():
    run("body_matmul_pull", 2)


def test_matmul_pull_world3():
    run("body_matmul_pull", 3)


def test_data_parallel_world2():
    run("body_data_parallel", 2)


def test_data_parallel_bucket_world2():
    run("body_data_parallel_bucket", 2)
