"""one launch set of the attention kernels at the BASELINE config-2 shape (for ncu): python scripts/attn_one.py [fwd|bwd]"""
import os, sys, torch
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
for _ in range(3):
    L.check(lib.progen_local_attn_fwd_tc(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, n, w, h, dh, L.stream()))
    if 'bwd' in sys.argv:
        L.check(lib.progen_local_attn_bwd_tc(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), delta.data_ptr(), 0, 0, B, n, w, h, dh, L.stream()))
torch.cuda.synchronize()
