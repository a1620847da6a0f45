"""tcgen05 attention kernels vs the mma.sync kernels at the BASELINE config-2 shape (forward and backward)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progen_b200 import lib as L
L.require_device()
B, n, w, h, dh = 64, 1024, 256, 8, 64
T, I = B * n, h * dh
qkv = torch.randn(T, 3 * I, device='cuda').bfloat16()
out = torch.empty(T, I, device='cuda', dtype=torch.bfloat16)
dout = torch.randn(T, I, device='cuda').bfloat16()
dqkv = torch.empty_like(qkv)
lse = torch.empty(T, h, device='cuda'); delta = torch.empty(T, h, device='cuda')
lib = L.load()
def f_tc(): L.check(lib.progen_local_attn_fwd_tc(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, n, w, h, dh, L.stream()))
def f_mma(): L.check(lib.progen_local_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, n, w, h, dh, L.stream()))
def b_tc(): L.check(lib.progen_local_attn_bwd_tc(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), 0, 0, B, n, w, h, dh, L.stream()))
def b_mma(): L.check(lib.progen_local_attn_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), 0, 0, B, n, w, h, dh, L.stream()))
def t(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
flops = 4.0 * I * (w + (w + 1) / 2) * T
res = dict(fwd_tc_ms=round(t(f_tc), 4), fwd_mma_ms=round(t(f_mma), 4))
if 'fwd' not in sys.argv:
    res.update(bwd_tc_ms=round(t(b_tc), 4), bwd_mma_ms=round(t(b_mma), 4))
res['fwd_tc_tflops'] = round(flops / res['fwd_tc_ms'] / 1e9, 1)
print(json.dumps(res))
