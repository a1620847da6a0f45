"""Training / sampling glue with the reference's names (progen_transformer/utils.py)."""
import numpy as np
import torch


def exists(val):
    return val is not None


def confirm(question):
    while True:
        resp = input(f'{question} (y/n) ')
        if resp.lower() in ('y', 'n'):
            return resp.lower() == 'y'


def get_loss_fn(model, data_parallel=False):
    """`loss, grads = get_loss_fn(model)(params, key, data)` like utils.py:61-93 (value_and_grad of the batched masked CE).
    With data_parallel the caller passes this rank's rows (progen_b200.parallel.shard_batch) and the gradient is the
    all-reduced global mean; see Trainer for the device-resident loop."""
    def loss_fn(params, key, data):
        return model.loss_and_grad(params, data)
    return loss_fn


def select_top_k(logits, k):
    """utils.py:97-100 — `>` the k-th value keeps k-1 entries, the rest become 0.0 (not -inf); returns (mask, filtered)"""
    kth = torch.topk(logits, k).values.min()
    mask = logits > kth
    return mask, torch.where(mask, logits, torch.zeros_like(logits))


def gumbel_noise(gen, shape, device):
    u = torch.rand(shape, generator=gen, device=device)
    return -torch.log(-torch.log(u + 1e-20) + 1e-20)


def sample(rng, fn, params, prime, length, top_k=None, add_bos=False, greedy=False):
    """Bug-compatible restatement of utils.py:106-135 over a device forward `fn(params, key, seq) -> logits`:
    full re-forward per generated token, Gumbel-max over the quirky top-k filter, the add_bos off-by-one (the first
    sampled id is ADDED to the last prime token), truncation after the second pad.  greedy=True zeroes the noise."""
    prime = np.asarray(prime).astype(np.int64)
    start_pos = prime.shape[-1]
    pad_right = length - prime.shape[-1]
    padding = (0, pad_right) if not add_bos else (1, pad_right - 1)
    seq = np.pad(prime, padding)
    gen = None
    if not greedy:
        gen = torch.Generator(device='cuda')
        gen.manual_seed(int(rng) if isinstance(rng, (int, np.integer)) else 0)
    seq_t = torch.as_tensor(seq, device='cuda')
    for curr_pos in range(start_pos, length):
        logits = fn(params, None, seq_t)[curr_pos - 1].float()
        noise = torch.zeros_like(logits) if greedy else gumbel_noise(gen, logits.shape, logits.device)
        if exists(top_k):
            mask, logits = select_top_k(logits, top_k)
            noise = noise * mask
        sampled = torch.argmax(logits + noise, dim=-1)
        seq_t[curr_pos] += sampled
    seq = seq_t.cpu().numpy()
    after_eos = np.cumsum(seq == 0) > 1
    return seq * ~after_eos
