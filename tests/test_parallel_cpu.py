"""Host-side data-parallel logic on CPU with the gloo backend, world_size 2 (the N>1 path of bench.py / train.py):
row sharding follows the reference's '(p b)' split with ragged batches (utils.py:78-91), and a SUM all-reduce of
per-rank gradients scaled by 1/global_rows equals the single-process gradient of the masked mean."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_covers_batch_like_reference_padding():
    from progen_b200.parallel import shard_rows
    for rows in range(1, 20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(rows, r, world) for r in range(world)]
            got = [i for a, b in spans for i in range(a, b)]
            assert got == list(range(rows))
            per = -(-rows // world)
            assert all(b - a <= per for a, b in spans)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from progen_b200 import parallel as PAR
    from oracle import progen_ref as O
    from oracle import progen_torch as T
    torch.set_num_threads(2)
    kwargs = dict(num_tokens=256, dim=32, seq_len=16, depth=2, window_size=8, global_mlp_depth=1, heads=2, dim_head=16)
    cfg = O.make_config(**kwargs)
    params = O.randomize_params(O.init_params(cfg, 1), 2)
    data = np.random.default_rng(3).integers(0, 256, (5, cfg['seq_len'] + 1)).astype(np.int64)     # ragged: 5 rows / 2 ranks
    local = PAR.shard_batch(data)
    assert PAR.world() == (rank, world)
    # per-rank: sum of per-row losses / global_rows  (what Engine.loss_and_grad(global_batch=...) computes)
    prm = T.to_torch(params, torch.float64, requires_grad=True)
    ids, labels = torch.as_tensor(local[:, :-1]), torch.as_tensor(local[:, 1:])
    loss = T.cross_entropy(T.forward(prm, ids, cfg), labels).sum() / data.shape[0]
    loss.backward()
    keys = sorted((m, k) for m, d in prm.items() for k in d)
    flat = torch.cat([prm[m][k].grad.reshape(-1) for m, k in keys])
    PAR.allreduce_sum_(flat, bucket_elems=1000)          # several buckets
    lt = PAR.allreduce_scalar_(loss.detach().clone())
    if rank == 0:
        ref_loss, ref = T.loss_and_grads(params, data, cfg)
        ref_flat = np.concatenate([ref[m][k].ravel() for m, k in keys])
        ret['loss_err'] = abs(float(lt) - ref_loss)
        ret['grad_err'] = float(np.abs(flat.numpy() - ref_flat).max())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_equals_single_process_gradient():
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret['loss_err'] < 1e-12 and ret['grad_err'] < 1e-12, dict(ret)


def _bucket_worker(rank, world, port, ret):
    """The Trainer's overlapped, bucketed gradient all-reduce on a stand-in engine (flat CPU buffer, gloo): every element of
    the flat buffer is reduced exactly once, whatever the bucket size, and the result is the sum over ranks."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from progen_b200.trainer import Trainer

    class FakeEngine:
        # layout like the real one: [embedding/head | layer 0 | layer 1 | ... | layer L-1 | small ndim<=1 section]
        def __init__(self, layers=5, per_layer=24, head=16, tail=8):
            self.layers, self.per_layer, self.head = layers, per_layer, head
            self.n_params_padded = head + layers * per_layer + tail
            self.grads = torch.arange(self.n_params_padded, dtype=torch.float32) * (rank + 1)

        def layer_grad_range(self, i):
            a = self.head + i * self.per_layer
            return a, a + self.per_layer

    ok = True
    for bucket_layers in (1, 2, 3, 7):
        tr = Trainer.__new__(Trainer)
        tr.eng = FakeEngine()
        tr.world, tr.rank = world, rank
        tr._works, tr._done, tr._bucket, tr._bucket_layers = [], [], None, bucket_layers
        for i in reversed(range(tr.eng.layers)):            # backward visits the layers from last to first
            tr._reduce_layer(i)
        tr._finish_allreduce()
        expect = torch.arange(tr.eng.n_params_padded, dtype=torch.float32) * sum(r + 1 for r in range(world))
        ok = ok and bool(torch.equal(tr.eng.grads, expect)) and tr._bucket is None and not tr._works
    ret[rank] = ok
    dist.destroy_process_group()


def test_trainer_bucketed_allreduce_covers_every_gradient_once():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, 29641, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
