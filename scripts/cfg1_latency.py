"""BASELINE configs[0]: ProGen dim=512 depth=2 seq_len=1024 window=256, ONE sequence forward (model.apply), device-timed.
python scripts/cfg1_latency.py  -> one JSON line (fp32 parity engine and bf16 tensor-core engine)"""
import json, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progen_b200 import ProGen

kw = dict(num_tokens=256, dim=512, seq_len=1024, depth=2, window_size=256)
seq = np.random.default_rng(0).integers(0, 256, 1024)
out = dict(config='ProGen dim=512 depth=2 seq_len=1024 window=256, single sequence forward (BASELINE configs[0])')
for mp in (False, True):
    model = ProGen(**kw, mixed_precision=mp)
    params = model.init(1)
    eng = model.engine
    model.apply(params, None, seq)
    ids = torch.as_tensor(seq.reshape(1, -1))
    eng.forward(ids)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        eng._forward_device()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    out['bf16' if mp else 'fp32'] = dict(forward_ms=round(ms, 4), tokens_per_sec=round(1024 / ms * 1e3))
print(json.dumps(out))
