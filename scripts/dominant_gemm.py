"""Launch only the kernels bench.py's `roofline` object times (bench.time_step_kernels: attention forward / backward, FF
proj_in weight gradient, FF proj_in + GLU, FF proj_out dgrad + GLU backward; config-2 shapes) a few times.

Meant to sit under `ncu --set full -k regex:<kernel> -c N --profile-from-start off` so the capture holds exactly the
kernels whose roofline bench.py reports; prints the CUDA-event timings when run bare."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import bench
from progen_b200 import ProGen


def main():
    cfgd = bench.CONFIGS[os.environ.get('CONFIG', 'cfg2')]
    kw = cfgd['kwargs']
    B = int(os.environ.get('BATCH', cfgd['batch']))
    model = ProGen(**kw, mixed_precision=True)
    tr = model.trainer(model.init(1234))
    eng = model.engine
    batch = bench.synthetic_batches(1, B, kw['seq_len'], 42)[0].cuda()
    eng.ensure_batch(B)
    eng.tok.copy_(batch[:, :-1].reshape(-1)); eng.labels.copy_(batch[:, 1:].reshape(-1))
    tr.step_resident(global_batch=B)           # populates the saved activations the GEMM reads
    torch.cuda.synchronize()
    torch.cuda.profiler.start()                # ncu --profile-from-start off: only the launches below are visible
    res = bench.time_step_kernels(eng, kw, iters=int(os.environ.get('ITERS', '5')))
    torch.cuda.profiler.stop()
    print('DOMINANT ' + json.dumps(res))


if __name__ == '__main__':
    main()
