"""Where does the end-to-end leg lose time at N > 1?  Same captured step graph, K steps per variant, device-timed (max over ranks):
A: inputs copied device->device, step_resident      B: eng.load_batch(pinned host) + step_resident
C: Trainer.step(pinned host)                         D: C + non-blocking loss copy to pinned host
usage: torchrun --nproc-per-node N scripts/e2e_diag.py"""
import json, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from progen_b200 import ProGen, lib as L
rank, local, world = int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
cfgd = bench.CONFIGS['cfg2']; kw = cfgd['kwargs']; B = cfgd['batch']; n = kw['seq_len']
model = ProGen(**kw, mixed_precision=True)
tr = model.trainer(model.init(1234)); eng = model.engine
K, W = 8, 3
batches = bench.synthetic_batches(W + K, B, n, 42 + rank)
dev = [b.cuda() for b in batches]
for i in range(W):
    eng.ensure_batch(B)
    eng.tok.copy_(dev[i][:, :-1].reshape(-1)); eng.labels.copy_(dev[i][:, 1:].reshape(-1))
    tr.step_resident(global_batch=B * world)
tr.capture_graph(B, B * world)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
host_losses = torch.empty(K, dtype=torch.float32).pin_memory()
def run(name, f):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    t0 = time.perf_counter()
    e0.record()
    for j in range(K): f(j, W + j)
    t_enq = time.perf_counter() - t0
    e1.record(); barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / K, t_enq * 1e3 / K], device='cuda')
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return name, round(float(ms[0]), 3), round(float(ms[1]), 3)
def A(j, i):
    eng.tok.copy_(dev[i][:, :-1].reshape(-1)); eng.labels.copy_(dev[i][:, 1:].reshape(-1)); tr.step_resident(global_batch=B * world)
def Bv(j, i):
    eng.load_batch(batches[i]); tr.step_resident(global_batch=B * world)
def C(j, i):
    tr.step(batches[i])
def D(j, i):
    host_losses[j:j + 1].copy_(tr.step(batches[i]).reshape(1), non_blocking=True)
res = [run('A dev->dev + step_resident', A), run('B load_batch(host) + step_resident', Bv), run('C Trainer.step(host)', C), run('D C + async loss copy', D), run('A again', A)]
if rank == 0:
    print(json.dumps(dict(world=world, ms_per_step_and_host_enqueue_ms=res)))
tr._graph = None
if world > 1:
    dist.barrier(); dist.destroy_process_group()
