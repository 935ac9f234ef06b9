"""pytest plugin for tests/test_reference_suite_cpu.py: routes ``vk.GPU()`` and the generator to the
NumPy stand-in device whose kernels are the CPU oracle (tests/fake_device.py).  Test infrastructure."""
import fake_device
import vulkpy_b200.vkarray as vkarray
from vulkpy_b200 import _backend as _b

vkarray.createGPU = lambda idx, priority: fake_device.FakeDevice()
_b.Xoshiro128pp = fake_device.FakeRng
