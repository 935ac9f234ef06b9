"""
Regenerates the fixtures under tests/golden/.  Run in the build container, where
/root/reference is mounted:  python tests/golden/make_golden.py

* shader_names.json      -- base names of the reference's compute shaders
                            (/root/reference/vulkpy/shader/*.comp; same list as setup.py:11-48).
                            The C ABI must resolve every one of them (tests/test_abi.py).
* reference_vectors.json -- value-level vectors PUBLISHED by the reference: the
                            Xoshiro128pp(seed=0) docstring (vulkpy/random.py:12-24) and the
                            arithmetic example (vulkpy/__init__.py:19-24).  Parsed from the
                            reference files so that a typo here cannot go unnoticed.
* prng_streams.npz       -- streams produced by oracle/vulkpy_oracle.py AFTER it reproduced the
                            docstring vectors above; they pin lane layout, tail chunks and state
                            persistence for the CUDA kernels (regression fixture, not an
                            independent pin).
The reference cannot be executed here (needs libvulkan + glslc + an ICD), so no fixture comes
from running it.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import vulkpy_oracle as orc  # noqa: E402


def main():
    names = sorted(f[:-5] for f in os.listdir(os.path.join(REF, "vulkpy", "shader")) if f.endswith(".comp"))
    with open(os.path.join(HERE, "shader_names.json"), "w") as f:
        json.dump(names, f, indent=0)

    doc = open(os.path.join(REF, "vulkpy", "random.py")).read()
    m = re.search(r"r\.random\(shape=\(3,\)\)\)\n\[([^\]]+)\]", doc)
    uniform = [float(x) for x in m.group(1).split()]
    m = re.search(r"r\.normal\(shape=\(3,\)\)\)\n\[([^\]]+)\]", doc)
    normal = [float(x) for x in m.group(1).split()]
    vec = {
        "source": "vulkpy/random.py:12-24 (Xoshiro128pp(gpu, seed=0), default size=64)",
        "random_3": uniform,
        "normal_3_after_random_3": normal,
    }
    with open(os.path.join(HERE, "reference_vectors.json"), "w") as f:
        json.dump(vec, f, indent=1)

    # the oracle must reproduce the published vectors before it may generate fixtures
    r = orc.Xoshiro128pp(64, 0)
    # numpy prints float32 with the shortest digits that round-trip, so the uniforms (pure integer
    # arithmetic + an exact float subtraction) must match bit for bit; the normals went through
    # the driver's log/sqrt/sin/cos, so they are compared to 2 ulp
    assert np.array_equal(r.random(3), np.asarray(uniform, dtype=np.float32)), "oracle != random.py:18"
    assert np.allclose(r.normal(3), np.asarray(normal, dtype=np.float32), rtol=3e-7, atol=0), "oracle != random.py:24"

    streams = {}
    for size, seed, calls in [
        (64, 0, [("u32", 3), ("f32", 17), ("u32", 64), ("u32", 65), ("f32", 1000), ("u32", 4096 + 5)]),
        (64, 1234, [("f32", 100000)]),
        (1, 7, [("u32", 50)]),
        (3, 7, [("u32", 10), ("f32", 10)]),
        (30, 99, [("u32", 1000)]),
        (256, 5, [("u32", 100), ("u32", 256 * 300 + 17), ("f32", 5)]),
        (1024, 42, [("f32", 1 << 16)]),
    ]:
        g = orc.Xoshiro128pp(size, seed)
        for ci, (kind, n) in enumerate(calls):
            out = g.randint(n) if kind == "u32" else g.random(n)
            streams[f"s{size}_seed{seed}_call{ci}_{kind}_{n}"] = out
        streams[f"s{size}_seed{seed}_final_state"] = g.state.copy()
    np.savez_compressed(os.path.join(HERE, "prng_streams.npz"), **streams)
    print("wrote", len(names), "shader names,", len(streams), "stream arrays")


if __name__ == "__main__":
    main()
