"""`ProGen` — the drop-in for `progen_transformer.ProGen` (reference progen.py:235-243).

Same constructor keywords (including the accepted-and-ignored `attn_dim`, `clamp_gate`), and an object with
`.init(rng, seq) -> params` and `.apply(params, rng, seq) -> logits`, where `params` is the haiku-shaped nested dict
`{module_path: {name: array}}` of the reference, so reference checkpoints and oracle parameters
interchange.  Everything below `.apply` runs as sm_100a kernels behind the C ABI (include/progen_b200.h).

Beyond the reference surface: `.apply` also accepts a batch (B, n); `.loss_and_grad(params, data)` is the fused
equivalent of `value_and_grad(get_loss_fn(model))` (utils.py:61-93); `.trainer(...)` owns device-resident training
state (parameters, Adam moments, apply_every accumulator) for the train.py loop.
"""
import numpy as np
import torch

from . import lib as L
from .engine import Engine, build_param_specs, P


def _trunc_normal(rng, shape, std):
    r = rng.standard_normal(int(np.prod(shape)))
    bad = np.abs(r) > 2.0
    while bad.any():
        r[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(r) > 2.0
    return (r.reshape(shape) * std).astype(np.float32)


class ProGen:
    def __init__(self, *, num_tokens, dim, seq_len, depth, window_size=256, global_mlp_depth=2, heads=8, dim_head=64,
                 ff_mult=4, ff_glu=True, attn_dim=None, clamp_gate=True, shift_tokens=True, mixed_precision=False,
                 mixed_precision_policy=None):
        # attn_dim / clamp_gate are accepted and ignored, exactly like the reference (progen.py:201-202)
        self.config = dict(num_tokens=num_tokens, dim=dim, seq_len=seq_len, depth=depth, window_size=window_size,
                           global_mlp_depth=global_mlp_depth, heads=heads, dim_head=dim_head, ff_mult=ff_mult,
                           ff_glu=ff_glu, attn_dim=attn_dim, clamp_gate=clamp_gate, shift_tokens=shift_tokens)
        assert seq_len % window_size == 0, 'sequence length must be divisible by the window size'   # progen.py:80
        self.mixed_precision = bool(mixed_precision)
        self._engine = None
        self._loaded = None

    # ---- engine (created lazily so that constructing a model does not need a GPU; using it does)
    @property
    def engine(self):
        if self._engine is None:
            self._engine = Engine(self.config, self.mixed_precision)
        return self._engine

    def param_shapes(self):
        specs, _, _ = build_param_specs(self.config)
        out = {}
        for s in specs:
            out.setdefault(s.module, {})[s.name] = s.shape
        return out

    # ---- reference surface
    def init(self, rng, seq=None):
        """`model.init(rng, seq)` (train.py:130-131).  `rng` may be an int seed or a key-like array.  Distributions are the
        haiku defaults the reference relies on: Linear w ~ TruncatedNormal(1/sqrt(fan_in)), b = 0; Embed ~
        TruncatedNormal(1); LayerNorm scale = 1; SGU spatial_weights ~ U(+-1e-3/n), spatial_biases = 1 (progen.py:172-176)."""
        seed = int(np.asarray(rng).ravel()[-1]) if not isinstance(rng, (int, np.integer)) else int(rng)
        g = np.random.default_rng(seed)
        n = self.config['seq_len']
        out = {}
        for module, names in self.param_shapes().items():
            out[module] = {}
            for name, shape in names.items():
                if name == 'embeddings':
                    a = _trunc_normal(g, shape, 1.0)
                elif name == 'w':
                    a = _trunc_normal(g, shape, shape[0] ** -0.5)
                elif name == 'spatial_weights':
                    a = g.uniform(-1e-3 / n, 1e-3 / n, shape).astype(np.float32)
                elif name in ('scale', 'spatial_biases'):
                    a = np.ones(shape, np.float32)
                else:
                    a = np.zeros(shape, np.float32)
                out[module][name] = a
        return out

    def _ensure_loaded(self, params):
        if self._loaded is not params:
            self.engine.load_params(params)
            self._loaded = params

    def apply(self, params, rng, seq):
        """`model.apply(params, rng, seq)`; `rng` is accepted and unused (no dropout — SURVEY Q11).
        seq: (n,) or (B, n) integers -> fp32 logits (n, V) or (B, n, V) as a torch CUDA tensor."""
        seq_t = torch.as_tensor(np.asarray(seq).astype(np.int64) if not isinstance(seq, torch.Tensor) else seq)
        single = seq_t.dim() == 1
        ids = seq_t.reshape(1, -1) if single else seq_t
        if ids.shape[-1] != self.config['seq_len']:
            raise L.ProgenError(f"sequence length {ids.shape[-1]} != constructor seq_len {self.config['seq_len']}")  # Q12
        self._ensure_loaded(params)
        logits = self.engine.forward(ids).view(ids.shape[0], ids.shape[1], -1)
        return logits[0].clone() if single else logits.clone()

    __call__ = apply

    def loss_and_grad(self, params, data):
        """Fused `loss, grads = value_and_grad(batched_loss_fn)(params, key, data)` (utils.py:61-76).
        data: (B, n+1) integers.  Returns (python float loss, haiku-shaped dict of numpy fp32 gradients)."""
        self._ensure_loaded(params)
        loss = self.engine.loss_and_grad(data)
        return float(loss.item()), self.engine.export_grads()

    def trainer(self, params, **optim_kwargs):
        from .trainer import Trainer
        return Trainer(self, params, **optim_kwargs)
