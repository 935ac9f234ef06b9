"""Drop-in check of the public API: every public class, method (with its argument names) and function
that the reference's Python layer defines (parsed with ``ast`` from /root/reference/vulkpy, nothing
imported or copied) exists in this package.  Skipped where the reference checkout is absent."""
import ast
import importlib
import inspect
import os

import pytest

REF = "/root/reference/vulkpy"
MODULES = {
    "vkarray": "vulkpy_b200.vkarray", "random": "vulkpy_b200.random", "util": "vulkpy_b200.util",
    "nn/core": "vulkpy_b200.nn.core", "nn/layers": "vulkpy_b200.nn.layers", "nn/losses": "vulkpy_b200.nn.losses",
    "nn/models": "vulkpy_b200.nn.models", "nn/optimizers": "vulkpy_b200.nn.optimizers",
    "nn/parameters": "vulkpy_b200.nn.parameters", "nn/initializers": "vulkpy_b200.nn.initializers",
    "nn/regularizers": "vulkpy_b200.nn.regularizers",
}
# deliberate differences: Q6 -- arrays need no waiting destructor on an in-order stream
ALLOWED = {"vkarray: _GPUArray.__del__"}


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_every_public_name_of_the_reference_exists():
    missing = []
    for rel, ours in MODULES.items():
        tree = ast.parse(open(os.path.join(REF, rel + ".py")).read())
        mod = importlib.import_module(ours)
        for node in tree.body:
            if isinstance(node, ast.ClassDef):
                if node.name.startswith("_") and node.name != "_GPUArray":
                    continue
                cls = getattr(mod, node.name, None)
                if cls is None:
                    missing.append(f"{rel}: class {node.name}")
                    continue
                for sub in node.body:
                    if not isinstance(sub, ast.FunctionDef):
                        continue
                    n = sub.name
                    if n.startswith("_") and not (n.startswith("__") and n.endswith("__")):
                        continue
                    if not hasattr(cls, n):
                        missing.append(f"{rel}: {node.name}.{n}")
                        continue
                    try:
                        have = list(inspect.signature(getattr(cls, n)).parameters)
                    except (TypeError, ValueError):
                        continue
                    if "args" in have or "kwargs" in have:
                        continue
                    for a in [x.arg for x in sub.args.args + sub.args.kwonlyargs]:
                        if a not in have:
                            missing.append(f"{rel}: {node.name}.{n}(... {a} ...)")
            elif isinstance(node, ast.FunctionDef) and not node.name.startswith("_"):
                if not hasattr(mod, node.name):
                    missing.append(f"{rel}: def {node.name}")
    assert set(missing) <= ALLOWED, sorted(set(missing) - ALLOWED)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")
def test_package_exports_match():
    import vulkpy_b200 as vk
    import vulkpy_b200.nn as nn
    for name in ("GPU", "U32Array", "Shape", "Array", "zeros", "random", "nn", "util"):     # vulkpy/__init__.py:26-28
        assert hasattr(vk, name), name
    tree = ast.parse(open(os.path.join(REF, "nn", "__init__.py")).read())
    exported = set()
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom):
            exported.update(a.asname or a.name for a in node.names if a.name != "*")
    lacking = sorted(n for n in exported if not n.startswith("_") and not hasattr(nn, n))
    assert not lacking, lacking
