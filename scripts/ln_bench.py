"""LayerNorm+shift backward (residual mode) at the config-2 shape, CUDA-event timed; PROGEN_LN_STREAM=0/1 picks the kernel."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from progen_b200 import lib as L


def main():
    T, d, n = 65536, int(os.environ.get('D', '512')), 1024
    dev = 'cuda'
    x = torch.randn(T, d, device=dev)
    scale = torch.randn(d, device=dev)
    mean = x.mean(1).contiguous()
    rstd = (x.var(1, unbiased=False) + 1e-5).rsqrt().contiguous()
    dy = torch.randn(T, d, device=dev).bfloat16()
    dres = torch.zeros(T, d, device=dev)
    lp = torch.empty(T, d, device=dev, dtype=torch.bfloat16)
    dscale = torch.zeros(d, device=dev)
    cs = torch.zeros(d, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    lib = L.load()

    def run():
        L.check(lib.progen_ln_shift_bwd(dy.data_ptr(), d, L.BF16, x.data_ptr(), d, L.F32, scale.data_ptr(), mean.data_ptr(),
                                        rstd.data_ptr(), dres.data_ptr(), lp.data_ptr(), d, dscale.data_ptr(), cs.data_ptr(),
                                        T, d, n, 1, 1, L.stream()))
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    bytes_ = T * d * (4 + 2 + 4 + 4 + 2)
    print(json.dumps(dict(kernel='ln_shift_bwd residual', stream=os.environ.get('PROGEN_LN_STREAM', '1'), T=T, d=d, ms=ms,
                          gbs=bytes_ / ms / 1e6)))


if __name__ == '__main__':
    main()
