"""fp32 CUDA-core GEMM (gemm_simt.cu) + every epilogue against a torch float64 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('a_mn,b_mn', [(False, True), (False, False), (True, True), (True, False)])
@pytest.mark.parametrize('epi', range(8))
def test_simt_gemm_epilogues(a_mn, b_mn, epi):
    from progen_b200 import lib as L
    from gemm_cases import run_case
    M, N, K = 200, 192, 136            # ragged M and K tails
    err, scale = run_case(L.BACKEND_SIMT, torch.float32, M, N, K, a_mn, b_mn, epi, seed=epi, seq_len=50 if epi == 1 else None)
    assert err <= 2e-5 * max(1.0, scale), (err, scale)


def test_simt_batched_causal_and_reduce():
    from progen_b200 import lib as L
    dev = 'cuda'
    g = torch.Generator(device=dev).manual_seed(1)
    B, n, C = 3, 160, 64
    Wm = torch.tril(torch.randn(n, n, generator=g, device=dev))
    X = torch.randn(B * n, C, generator=g, device=dev)
    out = torch.empty(B * n, C, device=dev)
    # out_b = Wm @ X_b   (A K-major shared, B MN-major batched, lower-causal K skipping)
    L.gemm(M=n, N=C, K=n, A=Wm, lda=n, B=X, ldb=C, b_mn=True, out=out, ldo=C, backend=L.BACKEND_SIMT, in_dtype=L.F32,
           batch=B, b_batch_rows=n, d_batch_rows=n, causal=1)
    ref = torch.einsum('mk,bkc->bmc', Wm.double(), X.view(B, n, C).double()).reshape(B * n, C)
    assert (out.double() - ref).abs().max().item() < 1e-4
    # out_b = Wm^T @ X_b   (A MN-major, upper-causal)
    L.gemm(M=n, N=C, K=n, A=Wm, lda=n, a_mn=True, B=X, ldb=C, b_mn=True, out=out, ldo=C, backend=L.BACKEND_SIMT,
           in_dtype=L.F32, batch=B, b_batch_rows=n, d_batch_rows=n, causal=2)
    ref = torch.einsum('km,bkc->bmc', Wm.double(), X.view(B, n, C).double()).reshape(B * n, C)
    assert (out.double() - ref).abs().max().item() < 1e-4
    # dW = tril(sum_b G_b @ X_b^T)  (batch_reduce + tril mask, atomic accumulate)
    G = torch.randn(B * n, C, generator=g, device=dev)
    dW = torch.zeros(n, n, device=dev)
    L.gemm(M=n, N=n, K=C, A=G, lda=C, B=X, ldb=C, out=dW, ldo=n, backend=L.BACKEND_SIMT, in_dtype=L.F32,
           epi=L.EPI_ACCUM, batch=B, a_batch_rows=n, b_batch_rows=n, batch_reduce=True, atomic=True, tril=True, tril_rows=n)
    ref = torch.tril(torch.einsum('bmc,bkc->mk', G.view(B, n, C).double(), X.view(B, n, C).double()))
    assert (dW.double() - ref).abs().max().item() < 1e-3
