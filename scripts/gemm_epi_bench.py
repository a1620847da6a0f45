"""Micro-benchmark of the tcgen05 GEMMs WITH the fused epilogues the model uses (config 2: d=512, hid=2048, T=65536).
Each line also gives the HBM floor of the epilogue traffic, so the gap to max(tensor time, HBM time) is visible."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from progen_b200 import lib as L

T, D, HID, I = 65536, 512, 2048, 512
dev = 'cuda'


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def report(name, ms, flops, bytes_):
    print(json.dumps(dict(name=name, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), epilogue_MB=round(bytes_ / 1e6, 1),
                          hbm_floor_ms=round(bytes_ / 6.5e9, 4))), flush=True)


def main():
    L.require_device()
    bf = torch.bfloat16
    y = torch.randn(T, D, device=dev).to(bf)
    w_in = torch.randn(D, 2 * HID, device=dev).to(bf) * 0.05          # (in, out): MN-major B
    b_in = torch.randn(2 * HID, device=dev)
    u = torch.empty(T, 2 * HID, device=dev, dtype=bf)
    hact = torch.empty(T, HID, device=dev, dtype=bf)
    glu = lambda: L.gemm(M=T, N=2 * HID, K=D, A=y, lda=D, B=w_in, ldb=2 * HID, b_mn=True, out=hact, ldo=HID, out2=u, ldo2=2 * HID,
                         bias=b_in, epi=L.EPI_GLU, backend=L.BACKEND_TC, in_dtype=L.BF16, out_dtype=L.BF16)
    report('ffin_fwd+GLU', timeit(glu), 2.0 * T * 2 * HID * D, T * (D * 2 + 2 * HID * 2 + HID * 2))

    w_out = torch.randn(HID, D, device=dev).to(bf) * 0.05             # (in, out)
    b_out = torch.randn(D, device=dev)
    x = torch.randn(T, D, device=dev)
    res = lambda: L.gemm(M=T, N=D, K=HID, A=hact, lda=HID, B=w_out, ldb=D, b_mn=True, out=x, ldo=D, bias=b_out,
                         epi=L.EPI_RESIDUAL, backend=L.BACKEND_TC, in_dtype=L.BF16, out_dtype=L.F32)
    report('ffout_fwd+RESIDUAL', timeit(res), 2.0 * T * D * HID, T * (HID * 2 + D * 8))
    att = torch.randn(T, I, device=dev).to(bf)
    w_o = torch.randn(I, D, device=dev).to(bf) * 0.05
    res2 = lambda: L.gemm(M=T, N=D, K=I, A=att, lda=I, B=w_o, ldb=D, b_mn=True, out=x, ldo=D, bias=b_out,
                          epi=L.EPI_RESIDUAL, backend=L.BACKEND_TC, in_dtype=L.BF16, out_dtype=L.F32)
    report('attn_out+RESIDUAL', timeit(res2), 2.0 * T * D * I, T * (I * 2 + D * 8))

    # dgrad of proj_out with the GLU backward fused: dh = dy @ w_out^T (B = w_out is [N=HID, K=D] K-major), du = glu'(u) * dh
    dy = torch.randn(T, D, device=dev).to(bf)
    du = torch.empty(T, 2 * HID, device=dev, dtype=bf)
    glub = lambda: L.gemm(M=T, N=HID, K=D, A=dy, lda=D, B=w_out, ldb=D, b_mn=False, out=du, ldo=2 * HID, aux=u, ldaux=2 * HID,
                          epi=L.EPI_GLU_BWD, backend=L.BACKEND_TC, in_dtype=L.BF16, out_dtype=L.BF16)
    report('ffout_dgrad+GLU_BWD', timeit(glub), 2.0 * T * HID * D, T * (D * 2 + 2 * HID * 2 * 2))

    w_qkv = torch.randn(D, 3 * I, device=dev).to(bf) * 0.05
    qkv = torch.empty(T, 3 * I, device=dev, dtype=bf)
    n, dh = 1024, 64
    pos = torch.arange(n, device=dev, dtype=torch.float32)[:, None] * (10000 ** (-torch.arange(0, dh, 2, device=dev) / dh))[None]
    sin, cos = pos.sin().contiguous(), pos.cos().contiguous()
    rot = lambda: L.gemm(M=T, N=3 * I, K=D, A=y, lda=D, B=w_qkv, ldb=3 * I, b_mn=True, out=qkv, ldo=3 * I, epi=L.EPI_ROTARY,
                         rot_sin=sin, rot_cos=cos, seq_len=n, dim_head=dh, backend=L.BACKEND_TC, in_dtype=L.BF16, out_dtype=L.BF16)
    report('qkv_fwd+ROTARY', timeit(rot), 2.0 * T * 3 * I * D, T * (D * 2 + 3 * I * 2))


if __name__ == '__main__':
    main()
