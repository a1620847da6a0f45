"""BASELINE config 5: sample.py-style autoregressive decode, seq_len=1024, prime '[Tax=Mammalia] #', top_k=25, add_bos.
Reports generated tokens / device second for the KV-cached kernels (fp32 and bf16 weights) and, beside it, the CPU oracle
running the reference algorithm (full re-forward per token) on a bounded sample of tokens."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progen_b200 import ProGen
from progen_b200.decode import Decoder
from progen_b200.data import encode_tokens

kw = dict(num_tokens=256, dim=512, seq_len=1024, depth=12, heads=8, dim_head=64, window_size=256, global_mlp_depth=2)
model = ProGen(**kw)
params = model.init(1234)
prime = np.array(encode_tokens('[Tax=Mammalia] #'), dtype=np.uint16)
out = dict(workload='config 5: decode seq_len=1024, d=512 depth=12, prime 16 chars, top_k=25, add_bos, greedy')
for name, dt in (('fp32_weights', torch.float32), ('bf16_weights', torch.bfloat16)):
    dec = Decoder(model.config, params, weights_dtype=dt)
    dec.sample(prime, top_k=25, add_bos=True, greedy=True)            # builds the graph
    ids, steps, secs = dec.sample(prime, top_k=25, add_bos=True, greedy=True)
    out[name] = dict(tokens=steps, seconds=secs, tokens_per_sec=steps / secs, us_per_token=1e6 * secs / steps)
if '--cpu' in sys.argv:
    from oracle import progen_ref as O
    from oracle import progen_torch as T
    cfg = O.make_config(**kw)
    prm = T.to_torch(params, torch.float32)
    seq = torch.zeros(1, 1024, dtype=torch.long)
    ntok = 3
    with torch.no_grad():
        T.forward(prm, seq, cfg)
        t0 = time.perf_counter()
        for _ in range(ntok):
            T.forward(prm, seq, cfg)
        dt_ = (time.perf_counter() - t0) / ntok
    out['cpu_reference_algorithm'] = dict(kind='port', cores=len(os.sched_getaffinity(0)), tokens_per_sec=1.0 / dt_,
                                          sample=f'{ntok} tokens, each a full 1024-token fp32 forward (utils.py:115-117)')
print(json.dumps(out))
