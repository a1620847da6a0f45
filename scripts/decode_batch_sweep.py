"""decode throughput vs batch (BASELINE configs[4] model, bf16 weights, stochastic sampler): one JSON line per batch size"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from progen_b200 import ProGen, lib as L
from progen_b200.decode import BatchDecoder
from progen_b200.data import encode_tokens
L.require_device()
kw = bench.CONFIGS['cfg5']['kwargs']
model = ProGen(**kw)
params = model.init(1234)
prime = np.array(encode_tokens('[Tax=Mammalia] #'), dtype=np.int64)
for B in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64]:
    dec = BatchDecoder(model.config, params, batch=B, weights_dtype=torch.bfloat16)
    pr = [prime] * B if B > 1 else prime
    dec.sample(pr, top_k=25, add_bos=True, greedy=False, seed=1)
    best = None
    for s in (2, 3):
        ids, gen, secs = dec.sample(pr, top_k=25, add_bos=True, greedy=False, seed=s)
        best = secs if best is None else min(best, secs)
    print(json.dumps(dict(batch=B, tokens_per_sec=round(gen / best, 1), ms_per_step=round(best / (gen / B) * 1e3, 4))), flush=True)
    del dec
    torch.cuda.empty_cache()
