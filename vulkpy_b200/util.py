"""
Utilities (:mod:`vulkpy_b200.util`; reference: vulkpy/util.py)

``enable_debug`` of the reference switches on Vulkan validation / API-dump layers
(util.py:19-55).  The CUDA analogue: debug logging plus a synchronise-and-check after every
kernel launch, so that a faulting kernel is reported at the call that enqueued it.
"""
from __future__ import annotations

import logging
import os

logger = logging.getLogger("vulkpy")

_debug_sync = False


def enable_debug(*, validation: bool = True, api_dump: bool = True):
    """
    Enable debug mode.

    Parameters
    ----------
    validation : bool, optional
        Synchronise and check for CUDA errors after every launch (reference: validation layer).
    api_dump : bool, optional
        Log every submitted operation at DEBUG level (reference: API dump layer).
    """
    global _debug_sync
    logging.basicConfig(level=logging.DEBUG)
    logger.setLevel(logging.DEBUG)
    logger.debug("Enable debug mode")
    if validation:
        _debug_sync = True
        os.environ["VULKPY_DEBUG_SYNC"] = "1"
        from . import _backend
        import ctypes
        for dev in range(_backend.device_count() if _has_device() else 0):
            _backend.createGPU(dev, 0.0).set_debug_sync(True)


def _has_device() -> bool:
    try:
        from . import _backend
        return _backend.device_count() > 0
    except RuntimeError:
        return False


def getShader(name: str) -> str:
    """Kernel name for a reference shader file name: ``getShader("add.spv") -> "add"``
    (reference returns the .spv path: util.py:58-72; the backend resolves both forms)."""
    return name[:-4] if name.endswith(".spv") else name
