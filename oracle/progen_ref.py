"""CPU oracle for the ProGen hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this module, and only as the checker / timed CPU baseline.  The product (`progen_b200/`) never imports it and
fails loudly when its CUDA library is missing.

What it is: a dependency-free NumPy restatement of the reference algorithm (lucidrains/progen @ 3054b29), each
function citing the reference lines it follows.  The arithmetic primitives live in un-vendored third-party
libraries (jax ^0.2.20, dm-haiku ^0.0.4, optax ^0.0.9 — `pyproject.toml:10-21`), none importable here, and the
reference ships no tests or golden vectors, so **parity is unpinned by the reference's own tests**.  The pin we
do have: `tests/golden/make_golden.py` executes the reference's *own* `progen_transformer/progen.py` and
`utils.py` source under a numpy shim of those libraries (`oracle/ref_shim/`) and this oracle must reproduce its
logits / loss / greedy samples to fp64 round-off (`tests/test_oracle_golden.py`).  Library defaults restated
(not visible in the reference tree): tanh-approximate GELU, LayerNorm eps=1e-5 with biased variance and scale
only, Linear `y = x @ w + b` with `(in, out)` weights, haiku module-path parameter names.
"""
import math
import numpy as np

ATTN_MASK_VALUE = -1e10            # progen.py:18
LN_EPS = 1e-5                      # hk.LayerNorm default

# ----------------------------------------------------------------------------------------------------------
# configuration (progen.py:188-203 constructor defaults)


def make_config(*, num_tokens, dim, seq_len, depth, window_size=256, global_mlp_depth=2, heads=8, dim_head=64,
                ff_mult=4, ff_glu=True, attn_dim=None, clamp_gate=True, shift_tokens=True):
    assert seq_len % window_size == 0, 'sequence length must be divisible by the window size'   # progen.py:80
    return dict(num_tokens=num_tokens, dim=dim, seq_len=seq_len, depth=depth, window_size=window_size,
                global_mlp_depth=global_mlp_depth, heads=heads, dim_head=dim_head, ff_mult=ff_mult,
                ff_glu=ff_glu, attn_dim=attn_dim, clamp_gate=clamp_gate, shift_tokens=shift_tokens)


def layer_kinds(cfg):
    """progen.py:210-212 — layer i is gMLP iff (depth - i) <= global_mlp_depth; GLU iff not gMLP and ff_glu."""
    kinds = []
    for i in range(cfg['depth']):
        use_gmlp = (cfg['depth'] - i) <= cfg['global_mlp_depth']
        kinds.append('sgu' if use_gmlp else ('glu' if cfg['ff_glu'] else 'gelu'))
    return kinds


P = 'pro_gen_base/~/'              # haiku module-path prefix


def _trunc_normal(rng, shape, std):
    r = rng.standard_normal(int(np.prod(shape)))
    bad = np.abs(r) > 2.0
    while bad.any():
        r[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(r) > 2.0
    return (r.reshape(shape) * std).astype(np.float32)


def init_params(cfg, seed=0):
    """Distribution-level restatement of `model.init` (train.py:130-131): hk.Linear w ~ TruncatedNormal(1/sqrt(in)),
    b = 0; hk.Embed ~ TruncatedNormal(1); LN scale = 1; SGU weights ~ U(+-eps/n), biases = 1 (progen.py:172-176).
    Returns the haiku-shaped nested dict {module_path: {name: float32 ndarray}}."""
    rng = np.random.default_rng(seed)
    d, V, n = cfg['dim'], cfg['num_tokens'], cfg['seq_len']
    inner = cfg['heads'] * cfg['dim_head']
    hid = d * cfg['ff_mult']
    prm = {}
    prm[P + 'embed'] = {'embeddings': _trunc_normal(rng, (V, d), 1.0)}
    for i, kind in enumerate(layer_kinds(cfg)):
        a = P + f'attn{i}/~/'
        prm[a + 'layer_norm'] = {'scale': np.ones(d, np.float32)}
        prm[a + 'linear'] = {'w': _trunc_normal(rng, (d, inner * 3), d ** -0.5)}
        prm[a + 'linear_1'] = {'w': _trunc_normal(rng, (inner, d), inner ** -0.5), 'b': np.zeros(d, np.float32)}
        f = P + f'ff{i}/~/'
        h_in = hid * 2 if kind == 'glu' else hid
        h_out = hid // 2 if kind == 'sgu' else hid
        prm[f + 'layer_norm'] = {'scale': np.ones(d, np.float32)}
        prm[f + 'linear'] = {'w': _trunc_normal(rng, (d, h_in), d ** -0.5), 'b': np.zeros(h_in, np.float32)}
        if kind == 'sgu':
            half = hid // 2
            prm[f + 'sgu/~/layer_norm'] = {'scale': np.ones(half, np.float32)}
            eps = 1e-3 / n
            prm[f + 'sgu'] = {'spatial_weights': rng.uniform(-eps, eps, (n, n)).astype(np.float32),
                              'spatial_biases': np.ones((n, 1), np.float32)}
            prm[f + 'sgu/~/linear'] = {'w': _trunc_normal(rng, (half, half), half ** -0.5),
                                       'b': np.zeros(half, np.float32)}
        prm[f + 'linear_1'] = {'w': _trunc_normal(rng, (h_out, d), h_out ** -0.5), 'b': np.zeros(d, np.float32)}
    prm[P + 'layer_norm'] = {'scale': np.ones(d, np.float32)}
    prm[P + 'linear'] = {'w': _trunc_normal(rng, (d, V), d ** -0.5), 'b': np.zeros(V, np.float32)}
    return prm


def randomize_params(prm, seed=1, scale=0.3):
    """Perturb every parameter (biases, LN scales, SGU matrices included) so parity tests exercise paths that the
    default init leaves at 0 / 1 / ~1e-6."""
    rng = np.random.default_rng(seed)
    out = {}
    for mod, d in prm.items():
        out[mod] = {}
        for name, a in d.items():
            if name == 'spatial_weights':
                n = a.shape[0]
                out[mod][name] = (rng.standard_normal(a.shape) * (1.0 / n)).astype(np.float32)
            elif name in ('b', 'scale', 'spatial_biases'):
                out[mod][name] = (a + scale * rng.standard_normal(a.shape)).astype(np.float32)
            else:
                out[mod][name] = a
    return out


def num_params(prm):
    return sum(a.size for d in prm.values() for a in d.values())

# ----------------------------------------------------------------------------------------------------------
# helpers (progen.py:22-46)


def layer_norm(x, scale):
    """hk.LayerNorm(axis=-1, create_scale=True, create_offset=False) — progen.py:22."""
    mean = x.mean(axis=-1, keepdims=True)
    var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
    return (x - mean) / np.sqrt(var + LN_EPS) * scale


def fixed_pos_embedding(n, dim, dtype):
    """progen.py:24-28 — every frequency repeated twice adjacently."""
    inv_freq = 1.0 / (10000 ** (np.arange(0, dim, 2) / dim))
    ang = np.arange(n)[:, None] * inv_freq[None, :]
    ang = np.repeat(ang, 2, axis=-1)
    return np.sin(ang).astype(dtype), np.cos(ang).astype(dtype)


def rotate_every_two(x):
    """progen.py:30-34 — (x0, x1) -> (-x1, x0) on adjacent pairs."""
    out = np.empty_like(x)
    out[..., 0::2] = -x[..., 1::2]
    out[..., 1::2] = x[..., 0::2]
    return out


def apply_rotary(x, sin, cos):
    """progen.py:36-41 — rot_dim == dim_head, nothing passes through."""
    return x * cos + rotate_every_two(x) * sin


def shift_tokens(x):
    """progen.py:43-46 — first half of channels takes the value of the previous position (zeros at t = 0)."""
    half = (x.shape[-1] + 1) // 2          # np.array_split gives the first chunk the extra element
    out = x.copy()
    out[1:, :half] = x[:-1, :half]
    out[0, :half] = 0
    return out


def gelu(x):
    """jax.nn.gelu default (approximate=True, tanh form) — progen.py:141,143."""
    return 0.5 * x * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def softmax(x):
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)

# ----------------------------------------------------------------------------------------------------------
# blocks


def local_attention(x, prm, i, cfg, sin, cos):
    """progen.py:73-103."""
    a = P + f'attn{i}/~/'
    n, h, w, dh = x.shape[0], cfg['heads'], cfg['window_size'], cfg['dim_head']
    x = layer_norm(x, prm[a + 'layer_norm']['scale'])
    if cfg['shift_tokens']:
        x = shift_tokens(x)
    qkv = x @ prm[a + 'linear']['w']                                   # no bias (:70)
    q, k, v = np.split(qkv, 3, axis=-1)
    q, k, v = (t.reshape(n, h, dh).transpose(1, 0, 2) for t in (q, k, v))      # n (h d) -> h n d
    q, k, v = (apply_rotary(t, sin, cos) for t in (q, k, v))                   # rotary on q, k AND v (:87)
    W = n // w
    q, k, v = (t.reshape(h, W, w, dh) for t in (q, k, v))
    # one zero window in front, then every window sees (previous, current): 2w keys (:90-91)
    k, v = (np.concatenate((np.zeros_like(t[:, :1]), t), axis=1) for t in (k, v))
    k, v = (np.concatenate((t[:, :-1], t[:, 1:]), axis=2) for t in (k, v))
    sim = np.einsum('hwid,hwjd->hwij', q, k) * (dh ** -0.5)
    mask = np.tril(np.ones((w, 2 * w)), w).astype(bool)                        # row i sees cols j <= i + w
    sim = np.where(mask, sim, ATTN_MASK_VALUE)
    attn = softmax(sim)
    out = np.einsum('hwij,hwjd->hwid', attn, v)
    out = out.reshape(h, n, dh).transpose(1, 0, 2).reshape(n, h * dh)          # h w n d -> (w n) (h d)
    return out @ prm[a + 'linear_1']['w'] + prm[a + 'linear_1']['b']


def sgu(x, prm, i, cfg):
    """progen.py:166-185."""
    f = P + f'ff{i}/~/'
    n = cfg['seq_len']
    x, gate = np.split(x, 2, axis=-1)
    gate = layer_norm(gate, prm[f + 'sgu/~/layer_norm']['scale'])
    wts = prm[f + 'sgu']['spatial_weights'] * np.tril(np.ones((n, n)))
    gate = wts.astype(x.dtype) @ gate                                          # einsum 'n d, m n -> m d'
    gate = gate + prm[f + 'sgu']['spatial_biases']
    x = x * gate
    return x @ prm[f + 'sgu/~/linear']['w'] + prm[f + 'sgu/~/linear']['b']


def feed_forward(x, prm, i, cfg, kind):
    """progen.py:131-149."""
    f = P + f'ff{i}/~/'
    x = layer_norm(x, prm[f + 'layer_norm']['scale'])
    if cfg['shift_tokens']:
        x = shift_tokens(x)
    x = x @ prm[f + 'linear']['w'] + prm[f + 'linear']['b']
    if kind == 'glu':
        x, gate = np.split(x, 2, axis=-1)                                      # first half value, second gate
        x = x * gelu(gate)
    else:
        x = gelu(x)
    if kind == 'sgu':
        x = sgu(x, prm, i, cfg)
    return x @ prm[f + 'linear_1']['w'] + prm[f + 'linear_1']['b']


def forward(params, seq, cfg, dtype=np.float64):
    """`model.apply(params, rng, seq)` for ONE sequence — progen.py:224-233.  seq: (n,) ints -> (n, num_tokens)."""
    prm = {m: {k: v.astype(dtype) for k, v in d.items()} for m, d in params.items()}
    # jax clamps out-of-range gather indices (hk.Embed -> embeddings[ids]); reachable through the add_bos quirk (Q5)
    seq = np.clip(np.asarray(seq).astype(np.int64), 0, cfg['num_tokens'] - 1)
    n = seq.shape[0]
    x = prm[P + 'embed']['embeddings'][seq]
    sin, cos = fixed_pos_embedding(n, cfg['dim_head'], dtype)
    for i, kind in enumerate(layer_kinds(cfg)):
        x = x + local_attention(x, prm, i, cfg, sin, cos)
        x = x + feed_forward(x, prm, i, cfg, kind)
    x = layer_norm(x, prm[P + 'layer_norm']['scale'])
    return x @ prm[P + 'linear']['w'] + prm[P + 'linear']['b']

# ----------------------------------------------------------------------------------------------------------
# loss (utils.py:42-66)


def loss_mask(targets, ignore_index=0):
    """utils.py:54-56 — non-pad positions plus the FIRST pad (it doubles as end-of-string)."""
    mask = targets != ignore_index
    eos = np.cumsum(~mask, axis=-1) == 1
    return mask | (eos & ~mask)


def cross_entropy(logits, targets, ignore_index=0):
    """utils.py:45-59."""
    s = logits - logits.max(axis=-1, keepdims=True)
    logp = s - np.log(np.exp(s).sum(axis=-1, keepdims=True))
    nll = np.take_along_axis(logp, np.asarray(targets).astype(np.int64)[..., None], axis=-1)[..., 0]
    mask = loss_mask(np.asarray(targets), ignore_index)
    return -(nll * mask).sum(axis=-1) / mask.sum(axis=-1)


def batch_loss(params, data, cfg, dtype=np.float64):
    """utils.py:61-76 (non-data-parallel branch): data (B, n+1) -> mean over rows of CE(apply(data[:-1]), data[1:])."""
    vals = [cross_entropy(forward(params, row[:-1], cfg, dtype), row[1:]) for row in np.asarray(data)]
    return float(np.mean(vals))

# ----------------------------------------------------------------------------------------------------------
# sampler (utils.py:97-135), greedy limit: gumbel noise == 0


def select_top_k(logits, k):
    """utils.py:97-100 — `>` the k-th value keeps k-1 entries; the rest become 0.0, not -inf."""
    kth = np.sort(logits)[-k]
    mask = logits > kth
    return mask, np.where(mask, logits, 0.0)


def sample_greedy(params, prime, length, cfg, top_k=25, add_bos=False, dtype=np.float64, apply_fn=None):
    """utils.py:106-135 with zero noise.  Reproduces the add_bos off-by-one (Q5): the first sampled id is ADDED to
    the last prime token.  `apply_fn(seq) -> logits` may be supplied to test another implementation's forward."""
    prime = np.asarray(prime).astype(np.int64)
    start_pos = prime.shape[-1]
    pad_right = length - prime.shape[-1]
    padding = (0, pad_right) if not add_bos else (1, pad_right - 1)
    seq = np.pad(prime, padding)
    fn = apply_fn or (lambda s: forward(params, s, cfg, dtype))
    for curr_pos in range(start_pos, length):
        logits = np.asarray(fn(seq))[curr_pos - 1].astype(np.float64)
        if top_k is not None:
            _, logits = select_top_k(logits, top_k)
        seq[curr_pos] += int(np.argmax(logits))
    after_eos = np.cumsum(seq == 0) > 1                                          # utils.py:132-133
    return seq * ~after_eos

# ----------------------------------------------------------------------------------------------------------
# tokenizer (data.py:76-88)


def encode_tokens(s):
    return [ord(c) + 1 for c in s]


def decode_tokens(tokens, offset=1):
    return ''.join('' if t < 0 else chr(t) for t in (np.asarray(tokens).astype(np.int16) - offset))

# ----------------------------------------------------------------------------------------------------------
# optimizer: optax.chain(clip_by_global_norm, adamw(mask = ndim > 1), apply_every) — train.py:115-121,189-190


def optim_init(params, every=4):
    z = lambda: {m: {k: np.zeros(v.shape, np.float64) for k, v in d.items()} for m, d in params.items()}
    return dict(count=0, mu=z(), nu=z(), acc=z(), every_count=0, every=every)


def optim_step(params, grads, st, lr=2e-4, wd=1e-3, max_norm=0.5, b1=0.9, b2=0.999, eps=1e-8):
    """One `optim.update` + `apply_updates` (train.py:189-190), optax ^0.0.9 semantics:
    clip_by_global_norm: g *= max_norm / max(norm, max_norm);
    scale_by_adam (eps_root = 0) with bias correction; add_decayed_weights(wd) on ndim > 1 leaves; scale(-lr);
    apply_every(k): accumulate updates, emit the sum on every k-th call, zeros otherwise."""
    gn = math.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for d in grads.values() for g in d.values()))
    clip = max_norm / max(gn, max_norm)
    st['count'] += 1
    t = st['count']
    emit = st['every_count'] == st['every'] - 1
    new = {}
    for m, d in params.items():
        new[m] = {}
        for k, p in d.items():
            g = grads[m][k].astype(np.float64) * clip
            mu = st['mu'][m][k] = b1 * st['mu'][m][k] + (1 - b1) * g
            nu = st['nu'][m][k] = b2 * st['nu'][m][k] + (1 - b2) * g * g
            u = (mu / (1 - b1 ** t)) / (np.sqrt(nu / (1 - b2 ** t)) + eps)
            if p.ndim > 1:
                u = u + wd * p.astype(np.float64)
            u = -lr * u
            acc = st['acc'][m][k] + u
            if emit:
                new[m][k] = (p.astype(np.float64) + acc).astype(p.dtype)
                st['acc'][m][k] = np.zeros_like(acc)
            else:
                new[m][k] = p
                st['acc'][m][k] = acc
    st['every_count'] = (st['every_count'] + 1) % st['every']
    return new, gn
