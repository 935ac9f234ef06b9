"""Drop-in shims for maintainers of the reference.

``vulkpy_b200.compat._vkarray`` has the surface of the reference's pybind11 extension
``vulkpy._vkarray`` (/root/reference/vulkpy/_vkarray.cc:756-898) over the C ABI of
``include/vulkpy_b200.h``: copied (or imported as) ``vulkpy/_vkarray.py`` it lets the reference's
unmodified ``vkarray.py`` / ``random.py`` / ``nn`` drive the CUDA kernels.
"""
