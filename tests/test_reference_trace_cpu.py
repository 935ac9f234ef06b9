"""The recorded reference-suite trace (tests/golden/ref_suite_trace.*) is self-consistent and current:
replayed over the NumPy stand-in device (oracle kernels) it reproduces itself bit for bit, and every
shader name in it is one the C ABI knows.  The GPU counterpart is tests/test_gpu_reference_trace.py."""
import numpy as np

import fake_device
import ref_trace
from vulkpy_b200 import _backend as _b


def test_trace_replays_bit_exactly_over_the_oracle_device():
    calls, arrays = ref_trace.load_trace()
    dev = fake_device.FakeDevice()

    def make_buffer(pre, dtype):
        b = fake_device.FakeBuffer(dev, pre.size, np.uint32 if dtype == "uint32" else np.float32)
        b.arr[:] = pre
        return b
    n_dev, n_rng = ref_trace.replay(dev, lambda d, size, seed: fake_device.FakeRng(d, "", "", size, seed), calls, arrays,
                                    exact=True, make_buffer=make_buffer)
    assert n_dev >= 450 and n_rng >= 15


def test_trace_covers_the_shader_families():
    calls, _ = ref_trace.load_trace()
    ops = {c["op"] for c in calls if c.get("method") == "submit"}
    assert all(op in _b.OPS for op in ops)
    assert len(ops) >= 100                                   # 121 shaders; the suite touches most of them
    tests = {c["test"].split("::")[0] for c in calls}
    assert tests == {"test_vulkpy.py", "test_random.py", "test_nn.py"}
