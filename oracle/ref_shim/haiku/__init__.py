"""Minimal numpy stand-in for dm-haiku (^0.0.4) — TEST INFRASTRUCTURE ONLY, see ../README.md.

Restates only library behaviour the reference relies on (progen.py:22,70-71,125-126,164,173-176,207,
219-222,236): module-path naming (children built in a parent's __init__ live under `parent/~/child`,
repeated names get `_1`, `_2`, CamelCase -> snake_case), `hk.Linear` (`x @ w + b`, w:(in,out),
TruncatedNormal(1/sqrt(in)), b zeros), `hk.Embed` (TruncatedNormal(1.0)), `hk.LayerNorm`
(eps 1e-5, biased variance), `hk.Sequential`, `hk.get_parameter`, `hk.transform`, `hk.PRNGSequence`.
"""
import re
import numpy as np
from . import initializers  # noqa: F401

_DTYPE = np.float64


class _Frame:
    def __init__(self, params, is_init, rng):
        self.params = params
        self.is_init = is_init
        self.rng = rng
        self.stack = []          # [(module, method_name)]
        self.counters = {}       # scope-prefix -> {name: count}


_frames = []


def _frame():
    assert _frames, "haiku shim: modules must be used inside hk.transform"
    return _frames[-1]


def _camel_to_snake(name):
    return re.sub(r"((?<=[a-z0-9])[A-Z]|(?!^)[A-Z](?=[a-z]))", r"_\1", name).lower()


class _ModuleMeta(type):
    def __new__(mcs, name, bases, ns):
        for k, v in list(ns.items()):
            if callable(v) and k == '__call__':
                ns[k] = _wrap_method(k, v)
        return super().__new__(mcs, name, bases, ns)

    def __call__(cls, *args, **kwargs):
        fr = _frame()
        obj = cls.__new__(cls)
        obj._parent_ctx = fr.stack[-1] if fr.stack else None
        fr.stack.append((obj, '__init__'))
        try:
            obj.__init__(*args, **kwargs)
        finally:
            fr.stack.pop()
        return obj


def _wrap_method(name, fn):
    def wrapped(self, *a, **k):
        fr = _frame()
        fr.stack.append((self, name))
        try:
            return fn(self, *a, **k)
        finally:
            fr.stack.pop()
    return wrapped


class Module(metaclass=_ModuleMeta):
    def __init__(self, name=None):
        fr = _frame()
        base = name if name is not None else _camel_to_snake(type(self).__name__)
        ctx = self._parent_ctx
        if ctx is None:
            prefix = ''
        else:
            parent, method = ctx
            prefix = parent.module_name + '/' + ('~/' if method == '__init__' else '')
        cnt = fr.counters.setdefault(prefix, {})
        n = cnt.get(base, 0)
        cnt[base] = n + 1
        uniq = base if n == 0 else f'{base}_{n}'
        self.module_name = prefix + uniq
        self.name = uniq


def get_parameter(name, shape, dtype=None, init=None):
    fr = _frame()
    mod = fr.stack[-1][0]
    bucket = fr.params.setdefault(mod.module_name, {}) if fr.is_init else fr.params[mod.module_name]
    if name not in bucket:
        assert fr.is_init, f'missing parameter {mod.module_name}/{name}'
        initializers._RNG[0] = fr.rng
        bucket[name] = np.asarray(init(tuple(shape), _DTYPE), dtype=_DTYPE)
    p = bucket[name]
    assert tuple(p.shape) == tuple(shape), (mod.module_name, name, p.shape, shape)
    return p


class Linear(Module):
    def __init__(self, output_size, with_bias=True, w_init=None, b_init=None, name=None):
        super().__init__(name=name)
        self.output_size = output_size
        self.with_bias = with_bias
        self.w_init = w_init
        self.b_init = b_init

    def __call__(self, x):
        in_size = x.shape[-1]
        w_init = self.w_init or initializers.TruncatedNormal(stddev=1.0 / np.sqrt(in_size))
        w = get_parameter('w', (in_size, self.output_size), init=w_init)
        out = np.dot(x, w)
        if self.with_bias:
            b = get_parameter('b', (self.output_size,), init=self.b_init or (lambda s, d: np.zeros(s, d)))
            out = out + b
        return out


class Embed(Module):
    def __init__(self, vocab_size, embed_dim, w_init=None, name=None):
        super().__init__(name=name)
        self.vocab_size, self.embed_dim, self.w_init = vocab_size, embed_dim, w_init

    def __call__(self, ids):
        emb = get_parameter('embeddings', (self.vocab_size, self.embed_dim),
                            init=self.w_init or initializers.TruncatedNormal(stddev=1.0))
        return emb[np.clip(np.asarray(ids).astype(np.int64), 0, self.vocab_size - 1)]     # jax gather clamps


class LayerNorm(Module):
    def __init__(self, axis, create_scale, create_offset, eps=1e-5, scale_init=None, offset_init=None, name=None):
        super().__init__(name=name)
        self.axis, self.create_scale, self.create_offset, self.eps = axis, create_scale, create_offset, eps

    def __call__(self, x):
        mean = np.mean(x, axis=self.axis, keepdims=True)
        var = np.var(x, axis=self.axis, keepdims=True)          # biased, as jnp.var
        shape = (x.shape[self.axis],)
        scale = get_parameter('scale', shape, init=lambda s, d: np.ones(s, d)) if self.create_scale else 1.0
        offset = get_parameter('offset', shape, init=lambda s, d: np.zeros(s, d)) if self.create_offset else 0.0
        inv = scale / np.sqrt(var + self.eps)
        return inv * (x - mean) + offset


class Sequential(Module):
    def __init__(self, layers, name=None):
        super().__init__(name=name)
        self.layers = tuple(layers)

    def __call__(self, x):
        for l in self.layers:
            x = l(x)
        return x


class Transformed:
    def __init__(self, f):
        self._f = f

    def init(self, rng, *args, **kwargs):
        seed = int(np.asarray(rng).ravel()[-1])
        fr = _Frame({}, True, np.random.default_rng(seed))
        _frames.append(fr)
        try:
            self._f(*args, **kwargs)
        finally:
            _frames.pop()
        return fr.params

    def apply(self, params, rng, *args, **kwargs):
        fr = _Frame(params, False, None)
        _frames.append(fr)
        try:
            return self._f(*args, **kwargs)
        finally:
            _frames.pop()


def transform(f):
    return Transformed(f)


class PRNGSequence:
    def __init__(self, seed):
        self.seed = seed

    def __iter__(self):
        return self

    def __next__(self):
        return np.array([0, self.seed], dtype=np.uint32)


class mixed_precision:  # namespace stand-in
    @staticmethod
    def set_policy(cls, policy):
        raise NotImplementedError
