"""Target of an `ncu --profile-from-start off` capture of the persistent decode kernel (BASELINE configs[4] model): the caches are
filled up to position P by a normal run, then ONE launch of STEPS positions is bracketed by cudaProfilerStart/Stop.
usage: ncu --set full --clock-control none --profile-from-start off -o OUT python scripts/decode_ncu.py B [P] [STEPS]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from progen_b200 import ProGen, lib as L
from progen_b200.decode import BatchDecoder
from progen_b200.data import encode_tokens
L.require_device()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
P = int(sys.argv[2]) if len(sys.argv) > 2 else 500
STEPS = int(sys.argv[3]) if len(sys.argv) > 3 else 8
kw = bench.CONFIGS['cfg5']['kwargs']
model = ProGen(**kw)
params = model.init(1234)
prime = np.array(encode_tokens('[Tax=Mammalia] #'), dtype=np.int64)
dec = BatchDecoder(model.config, params, batch=B, weights_dtype=torch.bfloat16)
dec.sample([prime] * B if B > 1 else prime, top_k=25, add_bos=True, greedy=False, seed=1)     # fills the caches (all positions)
torch.cuda.synchronize()
torch.cuda.profiler.start()
dec.run(P, STEPS)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('profiled', B, P, STEPS)
