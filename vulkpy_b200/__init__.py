"""
vulkpy_b200: the vulkpy array hot path on NVIDIA B200
=====================================================

Drop-in for ``import vulkpy as vk`` (reference: vulkpy/__init__.py): ``vk.GPU``, ``vk.Array``,
``vk.U32Array``, ``vk.Shape``, ``vk.zeros``, ``vk.random``, ``vk.nn``, ``vk.util``.  The Vulkan
extension and the 121 GLSL shaders of the reference are replaced by ``libvulkpy_b200.so``:
hand-written sm_100a CUDA kernels behind a C ABI (``include/vulkpy_b200.h``).

>>> import vulkpy_b200 as vk
>>> gpu = vk.GPU()
>>> a = vk.Array(gpu, data=[1, 2, 3])
>>> b = vk.Array(gpu, data=[3, 3, 3])
>>> print(a + b)
[4. 5. 6.]
"""
from .vkarray import GPU, U32Array, Shape, Array, zeros, fuse
from . import random
from . import nn
from . import util
from ._backend import pinned_empty, device_count

__version__ = "0.1.0"
