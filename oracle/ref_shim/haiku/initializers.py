"""hk.initializers stand-ins: distributions match the library; streams are numpy's (the reference's init is not
seed-reproducible anyway — utils.py:155 set_hardware_rng_)."""
import numpy as np

_RNG = [None]

class TruncatedNormal:
    def __init__(self, stddev=1.0, mean=0.0):
        self.stddev, self.mean = float(stddev), float(mean)
    def __call__(self, shape, dtype):
        n = int(np.prod(shape))
        r = _RNG[0].standard_normal(n)
        bad = np.abs(r) > 2.0
        while bad.any():                      # rejection sampling: N(0,1) truncated to [-2, 2]
            r[bad] = _RNG[0].standard_normal(int(bad.sum()))
            bad = np.abs(r) > 2.0
        r = r.reshape(shape)
        return (r * self.stddev + self.mean).astype(dtype)

class RandomUniform:
    def __init__(self, minval=0.0, maxval=1.0):
        self.minval, self.maxval = minval, maxval
    def __call__(self, shape, dtype):
        return _RNG[0].uniform(self.minval, self.maxval, size=shape).astype(dtype)
