import numpy as np

def stop_gradient(x):
    return x

def top_k(x, k):
    idx = np.argsort(-x, axis=-1, kind='stable')[..., :k]
    return np.take_along_axis(x, idx, axis=-1), idx

def rng_uniform(a, b, shape):
    raise NotImplementedError

def convert_element_type(x, dtype):
    return np.asarray(x, dtype=dtype)
