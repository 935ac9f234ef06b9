"""
``import vulkpy`` compatibility alias.

The reference package is called ``vulkpy`` (reference: vulkpy/__init__.py:26-28); code and tests
written against it (``import vulkpy as vk``, ``from vulkpy.util import enable_debug``,
``from vulkpy.nn.parameters import Parameter``) run unchanged on the B200 backend when this
directory is on ``sys.path``: every ``vulkpy[.x]`` module name is bound to the matching
``vulkpy_b200[.x]`` module object.
"""
import sys as _sys

import vulkpy_b200 as _impl

for _name, _mod in list(_sys.modules.items()):
    if _name == "vulkpy_b200" or _name.startswith("vulkpy_b200."):
        _sys.modules["vulkpy" + _name[len("vulkpy_b200"):]] = _mod
