"""tcgen05 attention forward vs the mma.sync kernel at the BASELINE config-2 shape."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progen_b200 import lib as L
L.require_device()
B, n, w, h, dh = 64, 1024, 256, 8, 64
T, I = B * n, h * dh
qkv = torch.randn(T, 3 * I, device='cuda').bfloat16()
out = torch.empty(T, I, device='cuda', dtype=torch.bfloat16)
lse = torch.empty(T, h, device='cuda')
lib = L.load()
def f_tc(): L.check(lib.progen_local_attn_fwd_tc(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, n, w, h, dh, L.stream()))
def f_mma(): L.check(lib.progen_local_attn_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, n, w, h, dh, L.stream()))
def t(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it
flops = 4.0 * I * (w + (w + 1) / 2) * T
a, b = t(f_tc), t(f_mma)
print(json.dumps(dict(tc_ms=round(a, 4), mma_ms=round(b, 4), tc_tflops=round(flops / a / 1e9, 1), mma_tflops=round(flops / b / 1e9, 1))))
