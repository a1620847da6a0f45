"""LayerNorm(+token shift) backward at a config's [T, d]: streaming kernel (ln_stream.cu) vs PROGEN_LN_STREAM=0 (row per warp).
usage: python scripts/ln_bwd_bench.py D [T]   -> one JSON line with ms and achieved GB/s (algorithmic bytes)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progen_b200 import lib as L
L.require_device()
d = int(sys.argv[1]); T = int(sys.argv[2]) if len(sys.argv) > 2 else 16 * 1024
n = 1024
x = torch.randn(T, d, device='cuda'); scale = torch.randn(d, device='cuda')
y = torch.empty(T, d, device='cuda', dtype=torch.bfloat16)
mean = torch.empty(T, device='cuda'); rstd = torch.empty(T, device='cuda')
lib = L.load()
L.check(lib.progen_ln_shift_fwd(x.data_ptr(), d, L.F32, scale.data_ptr(), y.data_ptr(), d, L.BF16, mean.data_ptr(), rstd.data_ptr(), T, d, n, 1, L.stream()))
dy = torch.randn(T, d, device='cuda').bfloat16(); dres = torch.randn(T, d, device='cuda')
lp = torch.empty(T, d, device='cuda', dtype=torch.bfloat16); ds = torch.zeros(d, device='cuda'); cs = torch.zeros(d, device='cuda')
def f(): L.check(lib.progen_ln_shift_bwd(dy.data_ptr(), d, L.BF16, x.data_ptr(), d, L.F32, scale.data_ptr(), mean.data_ptr(), rstd.data_ptr(), dres.data_ptr(), lp.data_ptr(), d, ds.data_ptr(), cs.data_ptr(), T, d, n, 1, 1, L.stream()))
for _ in range(3): f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): f()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
byt = T * d * (4 + 2 + 4 + 4 + 2)          # x, dy, dres in, dres out, low-precision copy out
print(json.dumps(dict(d=d, T=T, stream=os.environ.get('PROGEN_LN_STREAM', '1'), ms=round(ms, 4), gbps=round(byt / ms / 1e6, 1))))
