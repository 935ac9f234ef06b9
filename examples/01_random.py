"""Xoshiro128++ on the device: uniform [0, 1), normal, integers -- the same stream the reference's
C++ seeding and shaders produce (vulkpy/random.py:12-24 documents the first values for seed 0).

    python examples/01_random.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vulkpy_b200 as vk

gpu = vk.GPU()
r = vk.random.Xoshiro128pp(gpu, seed=0)
u = r.random(shape=(3,))
print("random(3)  :", u, " (reference docstring: [0.42977667 0.8235899  0.90622926])")
print("normal(3)  :", r.normal(shape=(3,)), " (reference docstring: [-2.3403292  0.7247794  0.7118352])")
print("randint(4) :", r.randint(shape=(4,)))
print("randrange  :", r.randrange(shape=(8,), low=10, high=20))

big = vk.random.Xoshiro128pp(gpu, size=1 << 20, seed=7)
x = big.normal(shape=(1 << 24,), mean=1.0, stddev=2.0)
print("normal(2^24, mean 1, stddev 2): mean %.4f  std %.4f" % (float(np.asarray(x.mean())[0]),
                                                               float(np.sqrt(np.asarray(((x - 1.0) ** 2.0).mean())[0]))))
idx = big.permutation(10)
print("permutation(10):", idx)
