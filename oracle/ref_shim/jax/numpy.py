"""`jax.numpy` -> numpy.  Arrays are a subclass whose in-place ops are out-of-place (jax arrays are immutable,
so `seq += x` in the reference (utils.py:129) rebinds with type promotion)."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import float32, float64, int32, uint16, uint8, ndarray  # noqa: F401

class JArray(_np.ndarray):
    def __iadd__(self, o):
        return _np.add(_np.asarray(self), _np.asarray(o)).view(JArray)
    def __imul__(self, o):
        return _np.multiply(_np.asarray(self), _np.asarray(o)).view(JArray)

def _wrap(x):
    return _np.asarray(x).view(JArray)

def array(x, dtype=None):
    return _wrap(_np.array(x, dtype=dtype))

def pad(x, pad_width, mode='constant', constant_values=0.):
    return _wrap(_np.pad(_np.asarray(x), pad_width, mode=mode, constant_values=constant_values))

def ones(shape, dtype=None):
    return _np.ones(shape, dtype=dtype if dtype is not None else _np.float64)

def eye(n, dtype=None):
    return _np.eye(n, dtype=dtype if dtype is not None else _np.float64)
