"""Typing helpers (reference: vulkpy/vktyping.py)."""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np

KeyType = Union[int, np.ndarray, slice, tuple]
ValueType = Union[int, float, np.ndarray, Tuple]


class Resource:
    """Marker base of everything a pending job may need to keep alive."""
    __slots__ = ()
