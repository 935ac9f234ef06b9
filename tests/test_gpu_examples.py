"""The three example programs run end to end (the reference's CI runs its examples the same way:
Dockerfile:49-57)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("script,args", [("00_arithmetic.py", ["--n", "1024"]), ("01_random.py", []),
                                         ("02_nn_iris.py", ["--nepoch", "60"]),
                                         ("02_nn_iris.py", ["--nepoch", "60", "--optimizer", "sgd"])])
def test_example_runs(script, args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", script), *args], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
