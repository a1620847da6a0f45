"""How much of a training step is the GPU idle between kernels?  torch.profiler (CUPTI) over a few warm steps:
prints the step wall time, the summed kernel time and the largest gaps between consecutive kernels."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity

import bench
from progen_b200 import ProGen


def main():
    cfgd = bench.CONFIGS[os.environ.get('CONFIG', 'cfg2')]
    kw = cfgd['kwargs']
    B = int(os.environ.get('BATCH', cfgd['batch']))
    model = ProGen(**kw, mixed_precision=True)
    tr = model.trainer(model.init(1234))
    eng = model.engine
    batches = [b.cuda() for b in bench.synthetic_batches(8, B, kw['seq_len'], 42)]
    eng.ensure_batch(B)

    def step(i):
        eng.tok.copy_(batches[i][:, :-1].reshape(-1)); eng.labels.copy_(batches[i][:, 1:].reshape(-1))
        tr.step_resident(global_batch=B)
    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for i in range(4, 8):
            step(i)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
    busy = sum(e.time_range.end - e.time_range.start for e in evs)
    gaps = []
    end = evs[0].time_range.end
    for prev, e in zip(evs, evs[1:]):
        g = e.time_range.start - end
        if g > 0:
            gaps.append((g, prev.name[:60], e.name[:60]))
        end = max(end, e.time_range.end)
    gaps.sort(reverse=True)
    print(json.dumps(dict(steps=4, span_ms=(t1 - t0) / 1e3, kernel_ms=busy / 1e3, idle_ms=sum(g for g, _, _ in gaps) / 1e3,
                          launches=len(evs), gaps_over_20us=sum(1 for g, _, _ in gaps if g > 20))))
    for g, a, b in gaps[:12]:
        print(f'{g:8.1f} us  after {a}  before {b}')


if __name__ == '__main__':
    main()
