"""`jax.nn` primitives with the library's documented defaults (jax ^0.2.20)."""
import numpy as np

def softmax(x, axis=-1):
    e = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return e / np.sum(e, axis=axis, keepdims=True)

def log_softmax(x, axis=-1):
    s = x - np.max(x, axis=axis, keepdims=True)
    return s - np.log(np.sum(np.exp(s), axis=axis, keepdims=True))

def gelu(x, approximate=True):
    # jax.nn.gelu default: approximate=True (tanh form)
    if approximate:
        c = np.sqrt(2.0 / np.pi).astype(x.dtype) if hasattr(x, 'dtype') else np.sqrt(2.0 / np.pi)
        return 0.5 * x * (1.0 + np.tanh(c * (x + 0.044715 * (x ** 3))))
    raise NotImplementedError('exact GELU is not the jax default and is not used by the reference')
