"""Per-phase time of the persistent decode kernel (BASELINE configs[4] model): clock64 at entry / exit of every grid barrier
of one position, on CTA 0 and the last CTA.  compute = barrier entry - previous barrier exit (this CTA's work in the phase),
wait = exit - entry (arrival of the slowest CTA + the barrier itself).
usage: python scripts/decode_profile.py [B ...]  (default 1 64)"""
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
depth, nsgu = kw['depth'], kw['global_mlp_depth']
names = []
for li in range(depth):
    names += ['qkv', 'attn', 'out', 'ffin'] + (['sgu', 'sguproj'] if li >= depth - nsgu else []) + ['ffout']
names += ['head', 'sample']
mhz = torch.cuda.clock_rate() if hasattr(torch.cuda, 'clock_rate') else 1900
for B in [int(a) for a in sys.argv[1:]] or [1, 64]:
    dec = BatchDecoder(model.config, params, batch=B, weights_dtype=torch.bfloat16)
    ids, gen, secs = dec.sample([prime] * B if B > 1 else prime, top_k=25, add_bos=True, greedy=False, seed=1)
    ids, gen, secs = dec.sample([prime] * B if B > 1 else prime, top_k=25, add_bos=True, greedy=False, seed=2)
    out = dict(batch=B, tokens_per_sec=round(gen / secs, 1), us_per_step=round(secs / (gen / B) * 1e6, 2))
    for pos in (300, 900):
        p = dec.profile_barriers(pos - 2, 3)                   # last step = position `pos` (caches hold the earlier run)
        ev = p.shape[1]
        assert ev == len(names), (ev, len(names))
        agg = {}
        for cta in (0, 1):
            t = p[cta].astype(np.float64)
            comp = np.concatenate([[0.0], t[1:, 0] - t[:-1, 1]])
            wait = t[:, 1] - t[:, 0]
            for nm, c, w in zip(names, comp, wait):
                a = agg.setdefault(nm, [0.0, 0.0, 0.0, 0.0, 0])
                a[2 * cta] += c; a[2 * cta + 1] += w
                a[4] += 1 if cta == 0 else 0
        tot = (p[0, -1, 1] - p[0, 0, 1])
        rows = {nm: dict(n=a[4], cta0_compute=round(a[0] / a[4]), cta0_wait=round(a[1] / a[4]), ctaN_compute=round(a[2] / a[4]),
                         ctaN_wait=round(a[3] / a[4])) for nm, a in agg.items()}
        # marks inside CTA 0's GEMV phases, relative to the barrier exit in front of the phase
        mk = dec.last_marks.astype(np.float64)
        magg = {}
        for i, nm in enumerate(names):
            if i == 0 or mk[i, 0] == 0:
                continue
            rel = [(mk[i, k] - p[0, i - 1, 1]) if mk[i, k] else -1.0 for k in range(8)]
            a = magg.setdefault(nm, [np.zeros(8), 0])
            a[0] += np.array(rel); a[1] += 1
        marks = {nm: [int(x) for x in (a[0] / a[1])] for nm, a in magg.items()}
        out[f'pos{pos}'] = dict(step_cycles=int(tot), per_phase_avg_cycles=rows,
                                marks_enter_staged_fma_final_prefetch_ln1_ln2=marks)
    print(json.dumps(out), flush=True)
