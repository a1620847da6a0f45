"""Shared helpers: rebuild the seeded inputs the golden files were generated from (tests/golden/make_golden.py)."""
import os
import numpy as np

from oracle import progen_ref as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CASES = {
    'tiny_glu_sgu': dict(num_tokens=256, dim=64, seq_len=32, depth=3, window_size=8, global_mlp_depth=1,
                         heads=2, dim_head=32),
    'tiny_gelu_sgu': dict(num_tokens=256, dim=64, seq_len=64, depth=2, window_size=16, global_mlp_depth=1,
                          heads=4, dim_head=16, ff_glu=False),
    'tiny_all_glu': dict(num_tokens=256, dim=128, seq_len=128, depth=2, window_size=64, global_mlp_depth=0,
                         heads=2, dim_head=64),
    'cfg1': dict(num_tokens=256, dim=512, seq_len=1024, depth=2, window_size=256),
}


def fingerprint(params):
    return float(sum(np.abs(a.astype(np.float64)).sum() for d in params.values() for a in d.values()))


def load_case(name):
    """-> (cfg, params, data, golden npz) with params/data regenerated from the stored seeds and verified."""
    g = np.load(os.path.join(GOLDEN_DIR, f'{name}.npz'))
    cfg = O.make_config(**CASES[name])
    pseed = int(g['param_seed'])
    params = O.randomize_params(O.init_params(cfg, pseed), pseed + 1000)
    assert abs(fingerprint(params) - float(g['param_fingerprint'])) < 1e-6 * float(g['param_fingerprint']), \
        'seeded parameters drifted from the ones the golden file was generated with'
    return cfg, params, g['data'], g
