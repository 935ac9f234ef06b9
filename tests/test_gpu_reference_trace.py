"""The reference's acceptance suite on the CUDA backend, in a form that travels to the GPU box.

``tests/golden/ref_suite_trace.{json,npz}`` hold every call the reference's own 233 tests
(/root/reference/test/test_vulkpy.py, test_random.py, test_nn.py -- its CI recipe, Dockerfile:27-33)
make below the Python layer, with the buffer contents before and after each call; they were
captured by ``tests/golden/record_ref_suite.py`` from a run in which all 233 reference tests passed,
i.e. the recorded answers are answers the reference's assertions accept.  Here every call is replayed
through the C ABI on the device (same shader name, same parameter bytes, same inputs, same buffer
aliasing) and the buffers are compared with the recorded ones: bit-exact for IEEE / integer / index
work, within the stated tolerance for transcendentals, reductions (summation order), contractions
and Box-Muller.  When the reference checkout IS present (development container), the suite itself
also runs in place on the CUDA device.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import ref_trace
from vulkpy_b200 import _backend as _b

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
F = np.float32

def test_replay_every_device_call_of_the_reference_suite(gpu):
    calls, arrays = ref_trace.load_trace()
    n_dev, n_rng = ref_trace.replay(gpu.gpu, lambda dev, size, seed: _b.Xoshiro128pp(dev, "", "", size, seed), calls, arrays)
    assert n_dev >= 450 and n_rng >= 15


REF = "/root/reference/test"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box): the trace above stands in")
def test_reference_suite_in_place_on_the_cuda_backend():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, env.get("PYTHONPATH", "")])
    env["PYTHONDONTWRITEBYTECODE"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", REF, "-q", "-p", "no:cacheprovider", "--rootdir", "/tmp"],
                       capture_output=True, text=True, env=env, cwd="/tmp", timeout=900)
    assert r.returncode == 0 and "failed" not in r.stdout, r.stdout[-2000:] + r.stderr[-1000:]
