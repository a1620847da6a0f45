"""numpy stand-in for the handful of `jax` names the reference imports (test infrastructure only)."""
import numpy as _np
from . import numpy, nn, lax, random, tree_util  # noqa: F401

def jit(f, *a, **k):
    return f

def vmap(f, in_axes=0, out_axes=0):
    def g(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = a.shape[ax]
        outs = []
        for i in range(n):
            outs.append(f(*[a if ax is None else _np.take(a, i, axis=ax) for a, ax in zip(args, axes)]))
        return _np.stack(outs, axis=out_axes)
    return g

def pmap(f, *a, **k):
    raise NotImplementedError("pmap is not exercised under the numpy shim")

def value_and_grad(f, *a, **k):
    raise NotImplementedError("value_and_grad is not available under the numpy shim")

def local_device_count():
    return 1

def tree_map(f, tree):
    return tree_util.tree_map(f, tree)
