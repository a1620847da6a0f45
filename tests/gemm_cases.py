"""Shared GEMM test driver: runs progen_gemm (either backend) against a plain torch fp32/fp64 reference of the same
matmul + epilogue.  Used by test_gpu_gemm_simt.py and test_gpu_gemm_tc.py."""
import math
import torch

from progen_b200 import lib as L


def gelu(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def gelu_grad(x):
    x = x.detach().clone().requires_grad_(True)
    gelu(x).sum().backward()
    return x.grad


def rotary_tables(n, dh, device):
    inv_freq = 1.0 / (10000 ** (torch.arange(0, dh, 2, dtype=torch.float64) / dh))
    ang = torch.arange(n, dtype=torch.float64)[:, None] * inv_freq[None, :]
    return torch.sin(ang).float().to(device).contiguous(), torch.cos(ang).float().to(device).contiguous()


def make_operand(rows_mn, K, mn_major, dtype, gen, device, scale=1.0):
    """Returns (logical [rows_mn, K] float64 view, stored tensor, ld)."""
    if mn_major:
        st = (torch.randn(K, rows_mn, generator=gen, device=device) * scale).to(dtype)
        return st.double().t(), st, rows_mn
    st = (torch.randn(rows_mn, K, generator=gen, device=device) * scale).to(dtype)
    return st.double(), st, K


def run_case(backend, dtype, M, N, K, a_mn, b_mn, epi, seed=0, split_k=1, seq_len=None, dim_head=32):
    """dtype: torch.float32 (SIMT) or torch.bfloat16.  Returns max abs error relative to the reference scale."""
    dev = 'cuda'
    gen = torch.Generator(device=dev).manual_seed(seed)
    A, A_st, lda = make_operand(M, K, a_mn, dtype, gen, dev)
    Bm, B_st, ldb = make_operand(N, K, b_mn, dtype, gen, dev, scale=K ** -0.5)
    acc = A @ Bm.t()                                         # float64 reference of the contraction
    in_dt = L.F32 if dtype == torch.float32 else L.BF16
    kw = dict(M=M, N=N, K=K, A=A_st, lda=lda, B=B_st, ldb=ldb, backend=backend, a_mn=a_mn, b_mn=b_mn, in_dtype=in_dt,
              out_dtype=in_dt, epi=epi, split_k=split_k)
    tol_scale = 1.0
    if epi == L.EPI_STORE:
        bias = torch.randn(N, generator=gen, device=dev)
        out = torch.empty(M, N, device=dev, dtype=dtype)
        L.gemm(out=out, ldo=N, bias=bias, **kw)
        ref = acc + bias.double()
        got = out.double()
    elif epi == L.EPI_ROTARY:
        n = seq_len or M
        sin, cos = rotary_tables(n, dim_head, dev)
        out = torch.empty(M, N, device=dev, dtype=dtype)
        L.gemm(out=out, ldo=N, rot_sin=sin, rot_cos=cos, seq_len=n, dim_head=dim_head, **kw)
        pos = torch.arange(M, device=dev) % n
        s = sin.double()[pos].repeat_interleave(2, dim=-1).repeat(1, N // dim_head)
        c = cos.double()[pos].repeat_interleave(2, dim=-1).repeat(1, N // dim_head)
        rot = torch.stack((-acc[:, 1::2], acc[:, 0::2]), dim=-1).flatten(-2)
        ref = acc * c + rot * s
        got = out.double()
    elif epi == L.EPI_RESIDUAL:
        bias = torch.randn(N, generator=gen, device=dev)
        res = torch.randn(M, N, generator=gen, device=dev)
        ref = res.double() + acc + bias.double()
        L.gemm(out=res, ldo=N, bias=bias, **kw)
        got = res.double()
    elif epi == L.EPI_GLU:
        bias = torch.randn(N, generator=gen, device=dev)
        pre = torch.empty(M, N, device=dev, dtype=dtype)
        out = torch.empty(M, N // 2, device=dev, dtype=dtype)
        L.gemm(out=out, ldo=N // 2, out2=pre, ldo2=N, bias=bias, **kw)
        p = acc + bias.double()
        ref = torch.cat((p, p[:, 0::2] * gelu(p[:, 1::2])), dim=1)
        got = torch.cat((pre.double(), out.double()), dim=1)
    elif epi == L.EPI_GELU:
        bias = torch.randn(N, generator=gen, device=dev)
        pre = torch.empty(M, N, device=dev, dtype=dtype)
        out = torch.empty(M, N, device=dev, dtype=dtype)
        L.gemm(out=out, ldo=N, out2=pre, ldo2=N, bias=bias, **kw)
        p = acc + bias.double()
        ref = torch.cat((p, gelu(p)), dim=1)
        got = torch.cat((pre.double(), out.double()), dim=1)
    elif epi == L.EPI_GLU_BWD:
        u = torch.randn(M, 2 * N, generator=gen, device=dev).to(dtype)
        out = torch.empty(M, 2 * N, device=dev, dtype=dtype)
        L.gemm(out=out, ldo=2 * N, aux=u, ldaux=2 * N, **kw)
        ud = u.double()
        val, gate = ud[:, 0::2], ud[:, 1::2]
        ref = torch.stack((acc * gelu(gate), acc * val * gelu_grad(gate)), dim=-1).flatten(-2)
        got = out.double()
    elif epi == L.EPI_GELU_BWD:
        u = torch.randn(M, N, generator=gen, device=dev).to(dtype)
        out = torch.empty(M, N, device=dev, dtype=dtype)
        L.gemm(out=out, ldo=N, aux=u, ldaux=N, **kw)
        ref = acc * gelu_grad(u.double())
        got = out.double()
    elif epi == L.EPI_ACCUM:
        out = torch.randn(M, N, generator=gen, device=dev)
        ref = out.double() + acc
        kw['out_dtype'] = L.F32
        L.gemm(out=out, ldo=N, atomic=(split_k > 1), **kw)
        got = out.double()
    else:
        raise ValueError(epi)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    return err, scale
