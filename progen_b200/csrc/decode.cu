// KV-cached autoregressive decode (BASELINE config 5): one new position per step instead of the reference's full
// re-forward per generated token (utils.py:115-117).  Equivalent because every mixing op of the model is causal:
// sliding-window attention (keys <= query), token shift (t-1), SGU (tril-masked spatial weights).
//
// Per layer the step keeps: rotated K/V rows of every position (the two windows a query can see are slices of it; window
// 0's zero look-back keys are added analytically), the previous position's LayerNorm halves for the two token shifts, and
// for gMLP layers the history of normalised gate rows.  All activations are fp32; weights are fp32 or bf16, stored
// TRANSPOSED ([out, in], K contiguous) so a warp streams one output row with 16-byte loads: the step is a pure
// weight-streaming (HBM/L2-bound) pass.  The whole step reads the position from device memory, so one captured CUDA
// graph replays for every token with no host round trip; sampling (top-k filter, Gumbel-max, reference quirks Q5/Q6)
// happens on the device.
#include "common.cuh"
#include "../../include/progen_b200.h"

namespace {

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = GV_THREADS / 32;

enum { DE_BIAS = 0, DE_ROTARY_CACHE = 1, DE_RESIDUAL = 2, DE_GLU = 3, DE_GELU = 4 };

struct GemvArgs {
  const void* wt;        // [N, K] (DE_GLU: [2H, K], rows j and j+H form one output)
  const float* x;        // [K]
  const float* bias;     // [N] or null
  float* out;            // see epilogues
  int N, K;
  // DE_ROTARY_CACHE
  float* kcache; float* vcache; const float* rot_sin; const float* rot_cos; int inner; int dim_head;
  const int* pos;
  // fused prologue: x <- shift_tokens(LayerNorm(x) * ln_scale) (progen.py:74-77 / 132-135 / 220); every block recomputes the
  // row statistics (K floats), block 0 updates the token-shift cache.  ln_prev is double-buffered by position parity:
  // read [pos & 1], write [(pos + 1) & 1], so no block races with block 0's update.
  const float* ln_scale; float* ln_prev; int ln_shift;
};

template <typename TW> __device__ __forceinline__ float dot_row(const TW* __restrict__ w, const float* __restrict__ xs, int K, int lane);
template <> __device__ __forceinline__ float dot_row<float>(const float* __restrict__ w, const float* __restrict__ xs, int K, int lane) {
  float s = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    const float4 a = *reinterpret_cast<const float4*>(w + k);
    s = fmaf(a.x, xs[k], s); s = fmaf(a.y, xs[k + 1], s); s = fmaf(a.z, xs[k + 2], s); s = fmaf(a.w, xs[k + 3], s);
  }
  return s;
}
template <> __device__ __forceinline__ float dot_row<bf16>(const bf16* __restrict__ w, const float* __restrict__ xs, int K, int lane) {
  float s = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    float v[8];
    load_vec<8>(w + k, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(v[i], xs[k + i], s);
  }
  return s;
}

// each warp produces two outputs: rows (2i, 2i+1), or for DE_GLU rows (i, i+H)
template <typename TW, int EPI>
__global__ void __launch_bounds__(GV_THREADS) decode_gemv_kernel(const GemvArgs a) {
  extern __shared__ float xs[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (a.ln_scale) {
    __shared__ float red[GV_WARPS];
    __shared__ float stat[2];
    float s = 0.f;
    for (int k = threadIdx.x; k < a.K; k += GV_THREADS) { const float v = a.x[k]; xs[k] = v; s += v; }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < GV_WARPS; ++i) t += red[i]; stat[0] = t / a.K; }
    __syncthreads();
    const float mean = stat[0];
    float q = 0.f;
    for (int k = threadIdx.x; k < a.K; k += GV_THREADS) { const float u = xs[k] - mean; q += u * u; }
    q = warp_sum(q);
    if (lane == 0) red[warp] = q;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < GV_WARPS; ++i) t += red[i]; stat[1] = rsqrtf(t / a.K + 1e-5f); }
    __syncthreads();
    const float rstd = stat[1];
    const int half = a.K >> 1;
    const int p = a.ln_shift ? *a.pos : 0;
    const float* prev_rd = a.ln_prev + (p & 1) * half;
    float* prev_wr = a.ln_prev + ((p + 1) & 1) * half;
    for (int k = threadIdx.x; k < a.K; k += GV_THREADS) {
      const float v = (xs[k] - mean) * rstd * a.ln_scale[k];
      if (a.ln_shift && k < half) {
        xs[k] = prev_rd[k];
        if (blockIdx.x == 0) prev_wr[k] = v;
      } else {
        xs[k] = v;
      }
    }
  } else {
    for (int k = threadIdx.x; k < a.K; k += GV_THREADS) xs[k] = a.x[k];
  }
  __syncthreads();
  const int pair = blockIdx.x * GV_WARPS + warp;
  const TW* W = reinterpret_cast<const TW*>(a.wt);
  int r0, r1;
  if (EPI == DE_GLU) { r0 = pair; r1 = pair + a.N; if (pair >= a.N) return; }
  else { r0 = 2 * pair; r1 = r0 + 1; if (r0 >= a.N) return; }
  float s0 = warp_sum(dot_row<TW>(W + (long long)r0 * a.K, xs, a.K, lane));
  float s1 = warp_sum(dot_row<TW>(W + (long long)r1 * a.K, xs, a.K, lane));
  if (lane != 0) return;
  if (a.bias) { s0 += a.bias[r0]; s1 += a.bias[r1]; }
  if (EPI == DE_BIAS) { a.out[r0] = s0; a.out[r1] = s1; }
  else if (EPI == DE_RESIDUAL) { a.out[r0] += s0; a.out[r1] += s1; }
  else if (EPI == DE_GELU) { a.out[r0] = gelu_tanh(s0); a.out[r1] = gelu_tanh(s1); }
  else if (EPI == DE_GLU) { a.out[r0] = s0 * gelu_tanh(s1); }
  else if (EPI == DE_ROTARY_CACHE) {
    const int p = *a.pos;
    const int j = (r0 % a.dim_head) >> 1;
    const float sn = a.rot_sin[p * (a.dim_head >> 1) + j], cs = a.rot_cos[p * (a.dim_head >> 1) + j];
    const float o0 = s0 * cs - s1 * sn, o1 = s1 * cs + s0 * sn;           // rotary on q, k AND v (progen.py:87)
    const int sec = r0 / a.inner, c = r0 % a.inner;
    float* dst = sec == 0 ? a.out + c : (sec == 1 ? a.kcache + (long long)p * a.inner + c : a.vcache + (long long)p * a.inner + c);
    dst[0] = o0; dst[1] = o1;
  }
}

// x = embed[clamp(seq[pos])]   (hk.Embed; out-of-range ids clamp like a jax gather — reachable through quirk Q5)
__global__ void decode_embed_kernel(const int* __restrict__ seq, const int* __restrict__ pos, const float* __restrict__ table,
                                    float* __restrict__ x, int d, int V) {
  int id = seq[*pos];
  id = id < 0 ? 0 : (id >= V ? V - 1 : id);
  for (int c = threadIdx.x; c < d; c += blockDim.x) x[c] = table[(long long)id * d + c];
}

// y = shift_tokens(LN(x) * scale) for ONE row: first half comes from `prev` (the previous position's LN output, zeros
// at position 0), which is then replaced by this row's first half.  shift == 0: plain LN.
__global__ void decode_ln_kernel(const float* __restrict__ x, const float* __restrict__ scale, float* __restrict__ prev,
                                 float* __restrict__ y, int d, int shift) {
  __shared__ float red[32];
  __shared__ float stat[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  float s = 0.f;
  for (int c = tid; c < d; c += blockDim.x) s += x[c];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < nw; ++i) t += red[i]; stat[0] = t / d; }
  __syncthreads();
  const float mean = stat[0];
  float q = 0.f;
  for (int c = tid; c < d; c += blockDim.x) { const float u = x[c] - mean; q += u * u; }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < nw; ++i) t += red[i]; stat[1] = rsqrtf(t / d + 1e-5f); }
  __syncthreads();
  const float rstd = stat[1];
  const int half = d >> 1;
  for (int c = tid; c < d; c += blockDim.x) {
    const float v = (x[c] - mean) * rstd * scale[c];
    if (shift && c < half) { y[c] = prev[c]; prev[c] = v; }
    else y[c] = v;
  }
}

// one block per head: softmax(q . K^T / sqrt(dh)) V over the visible keys of position p (progen.py:88-102)
__global__ void __launch_bounds__(256) decode_attn_kernel(const float* __restrict__ q, const float* __restrict__ kcache,
                                                          const float* __restrict__ vcache, const int* __restrict__ pos,
                                                          float* __restrict__ out, int w, int inner, int dh) {
  extern __shared__ float sm[];                 // [dh] q | [2w] probabilities | [8][dh] partial outputs
  float* sq = sm;
  float* sp = sm + dh;
  float* so = sp + 2 * w;
  __shared__ float red[8];
  __shared__ float bc[2];
  const int hh = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int p = *pos, win = p / w, i = p % w;
  const int key0 = win > 0 ? (win - 1) * w : win * w;        // first real key position
  const int nreal = (win > 0 ? w : 0) + i + 1;
  const float scale = 1.0f / sqrtf((float)dh);
  for (int c = tid; c < dh; c += 256) sq[c] = q[hh * dh + c];
  __syncthreads();
  float mx = win == 0 ? 0.f : -INFINITY;                     // zero look-back keys of window 0: logit 0 (quirk Q1)
  for (int j = tid; j < nreal; j += 256) {
    const float* kr = kcache + (long long)(key0 + j) * inner + hh * dh;
    float s = 0.f;
    for (int c = 0; c < dh; c += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(kr + c);
      s = fmaf(kv.x, sq[c], s); s = fmaf(kv.y, sq[c + 1], s); s = fmaf(kv.z, sq[c + 2], s); s = fmaf(kv.w, sq[c + 3], s);
    }
    s *= scale;
    sp[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (tid == 0) { float t = red[0]; for (int k = 1; k < 8; ++k) t = fmaxf(t, red[k]); bc[0] = t; }
  __syncthreads();
  mx = bc[0];
  float l = 0.f;
  for (int j = tid; j < nreal; j += 256) { const float e = expf(sp[j] - mx); sp[j] = e; l += e; }
  l = warp_sum(l);
  __syncthreads();
  if (lane == 0) red[warp] = l;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int k = 0; k < 8; ++k) t += red[k]; if (win == 0) t += (float)w * expf(-mx); bc[1] = t; }
  __syncthreads();
  const float inv = 1.f / bc[1];
  // out[c] = sum_j p_j v_j[c]: warp `warp` takes keys j = warp, warp+8, ...; lanes own channels
  for (int c = lane; c < dh; c += 32) {
    float acc = 0.f;
    for (int j = warp; j < nreal; j += 8) acc = fmaf(sp[j], vcache[(long long)(key0 + j) * inner + hh * dh + c], acc);
    so[warp * dh + c] = acc;
  }
  __syncthreads();
  for (int c = tid; c < dh; c += 256) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += so[k * dh + c];
    out[hh * dh + c] = t * inv;
  }
}

// SGU (progen.py:166-184) for one position p: gn = LN(gate) * scale -> history[p]; gate' = sum_{k<=p} W[p,k] history[k] + b[p];
// s = xs * gate'.  a = [xs | gate] (C channels each).  Launch 1: LN + history write.  Launch 2: the causal mix.
__global__ void decode_sgu_mix_kernel(const float* __restrict__ a, const float* __restrict__ hist, const float* __restrict__ wsp,
                                      const float* __restrict__ bsp, const int* __restrict__ pos, float* __restrict__ s_out,
                                      int C, int n) {
  const int p = *pos;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* wrow = wsp + (long long)p * n;
  float acc = 0.f;
  for (int k = 0; k <= p; ++k) acc = fmaf(__ldg(wrow + k), hist[(long long)k * C + c], acc);
  s_out[c] = a[c] * (acc + bsp[p]);
}

__global__ void decode_hist_write_kernel(const float* __restrict__ gn, const int* __restrict__ pos, float* __restrict__ hist, int C) {
  const int p = *pos;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) hist[(long long)p * C + c] = gn[c];
}

// Sampling exactly as utils.py:97-129: top-k filter keeps logits > (k-th largest), the rest become 0.0 and lose their
// noise; argmax(logits + gumbel); seq[pos + 1] += index (ADD, quirk Q5); finally pos += 1.  One block of V threads.
__global__ void decode_sample_kernel(const float* __restrict__ logits, const float* __restrict__ noise, int* __restrict__ seq,
                                     int* __restrict__ pos, float* __restrict__ logits_out, int V, int n, int top_k, int do_sample) {
  extern __shared__ float sv[];                 // [V] filtered + noise
  __shared__ float kth;
  __shared__ int best;
  const int t = threadIdx.x;
  const int p = *pos;
  const float v = t < V ? logits[t] : -INFINITY;
  if (logits_out && t < V) logits_out[(long long)p * V + t] = v;
  if (t < V) sv[t] = v;
  __syncthreads();
  if (do_sample && p + 1 < n) {
    float f = v, nz = (noise && t < V) ? noise[(long long)p * V + t] : 0.f;
    if (top_k > 0) {
      int gt = 0, ge = 0;
      for (int j = 0; j < V; ++j) { gt += sv[j] > v; ge += sv[j] >= v; }
      if (t < V && gt < top_k && top_k <= ge) kth = v;       // the k-th largest value (with multiplicity)
      __syncthreads();
      const bool keep = v > kth;
      f = keep ? v : 0.f;
      nz = keep ? nz : 0.f;
    }
    __syncthreads();
    if (t < V) sv[t] = f + nz;
    __syncthreads();
    if (t == 0) {
      int b = 0;
      float bv = sv[0];
      for (int j = 1; j < V; ++j) if (sv[j] > bv) { bv = sv[j]; b = j; }      // first maximal index, like argmax
      best = b;
      seq[p + 1] += b;
    }
  }
  __syncthreads();
  if (t == 0) *pos = p + 1;
}

template <typename TW>
int gemv(int epi, const GemvArgs& a, cudaStream_t s) {
  const int pairs = epi == DE_GLU ? a.N : (a.N + 1) / 2;
  const int grid = (pairs + GV_WARPS - 1) / GV_WARPS;
  const size_t sm = (size_t)a.K * sizeof(float);
  switch (epi) {
    case DE_BIAS: decode_gemv_kernel<TW, DE_BIAS><<<grid, GV_THREADS, sm, s>>>(a); break;
    case DE_ROTARY_CACHE: decode_gemv_kernel<TW, DE_ROTARY_CACHE><<<grid, GV_THREADS, sm, s>>>(a); break;
    case DE_RESIDUAL: decode_gemv_kernel<TW, DE_RESIDUAL><<<grid, GV_THREADS, sm, s>>>(a); break;
    case DE_GLU: decode_gemv_kernel<TW, DE_GLU><<<grid, GV_THREADS, sm, s>>>(a); break;
    case DE_GELU: decode_gemv_kernel<TW, DE_GELU><<<grid, GV_THREADS, sm, s>>>(a); break;
  }
  PG_LAUNCH_CHECK();
  return PROGEN_OK;
}

int gemv_dispatch(int wdtype, int epi, const GemvArgs& a, cudaStream_t s) {
  return wdtype == PG_BF16 ? gemv<bf16>(epi, a, s) : gemv<float>(epi, a, s);
}

}  // namespace

extern "C" {

// One decode step: consumes seq[*pos], advances every cache to position *pos, writes logits (and, when sampling,
// seq[*pos + 1] += sampled id), then *pos += 1.  Graph-capturable: no host-visible state changes.
int progen_decode_step(const progen_decode_t* m, int do_sample, void* stream) {
  PG_CHECK_ARG(m != nullptr && m->layers != nullptr && m->depth > 0);
  PG_CHECK_ARG(m->d % 8 == 0 && m->inner % 8 == 0 && m->hid % 8 == 0 && m->V <= 1024 && m->dim_head % 4 == 0);
  cudaStream_t s = (cudaStream_t)stream;
  const int d = m->d, I = m->inner, hid = m->hid, n = m->n;
  decode_embed_kernel<<<1, 256, 0, s>>>(m->seq, m->pos, m->embed, m->x, d, m->V);
  PG_LAUNCH_CHECK();
  for (int i = 0; i < m->depth; ++i) {
    const progen_decode_layer_t& L = m->layers[i];
    // ---- LocalAttention
    GemvArgs a{};
    a.wt = L.wqkv_t; a.x = m->x; a.bias = nullptr; a.out = m->q; a.N = 3 * I; a.K = d;
    a.ln_scale = L.ln1_scale; a.ln_prev = L.shift1; a.ln_shift = m->shift_tokens;       // LN + token shift fused in the prologue
    a.kcache = L.kcache; a.vcache = L.vcache; a.rot_sin = m->rot_sin; a.rot_cos = m->rot_cos; a.inner = I; a.dim_head = m->dim_head;
    a.pos = m->pos;
    int rc = gemv_dispatch(m->wdtype, DE_ROTARY_CACHE, a, s);
    if (rc) return rc;
    const size_t asm_bytes = (size_t)(m->dim_head + 2 * m->window + 8 * m->dim_head) * sizeof(float);
    decode_attn_kernel<<<m->heads, 256, asm_bytes, s>>>(m->q, L.kcache, L.vcache, m->pos, m->att, m->window, I, m->dim_head);
    PG_LAUNCH_CHECK();
    a = GemvArgs{};
    a.wt = L.wo_t; a.x = m->att; a.bias = L.bo; a.out = m->x; a.N = d; a.K = I;
    rc = gemv_dispatch(m->wdtype, DE_RESIDUAL, a, s);
    if (rc) return rc;
    // ---- FeedForward
    a = GemvArgs{};
    a.wt = L.win_t; a.x = m->x; a.bias = L.bin; a.out = m->u; a.K = d;
    a.ln_scale = L.ln2_scale; a.ln_prev = L.shift2; a.ln_shift = m->shift_tokens; a.pos = m->pos;
    const float* last = m->u;
    int last_k = hid;
    if (L.kind == 0) {            // GLU: rows [0,hid) value, [hid,2hid) gate
      a.N = hid;
      rc = gemv_dispatch(m->wdtype, DE_GLU, a, s);
    } else {
      a.N = hid;
      rc = gemv_dispatch(m->wdtype, DE_GELU, a, s);
    }
    if (rc) return rc;
    if (L.kind == 2) {            // SGU
      const int C = hid / 2;
      decode_ln_kernel<<<1, 256, 0, s>>>(m->u + C, L.sgu_ln_scale, nullptr, m->gn, C, 0);
      PG_LAUNCH_CHECK();
      decode_hist_write_kernel<<<(C + 255) / 256, 256, 0, s>>>(m->gn, m->pos, L.gn_hist, C);
      PG_LAUNCH_CHECK();
      decode_sgu_mix_kernel<<<(C + 127) / 128, 128, 0, s>>>(m->u, L.gn_hist, L.sgu_w, L.sgu_b, m->pos, m->sg, C, n);
      PG_LAUNCH_CHECK();
      a = GemvArgs{};
      a.wt = L.sgu_proj_t; a.x = m->sg; a.bias = L.sgu_proj_b; a.out = m->pj; a.N = C; a.K = C;
      rc = gemv_dispatch(m->wdtype, DE_BIAS, a, s);
      if (rc) return rc;
      last = m->pj; last_k = C;
    }
    a = GemvArgs{};
    a.wt = L.wout_t; a.x = last; a.bias = L.bout; a.out = m->x; a.N = d; a.K = last_k;
    rc = gemv_dispatch(m->wdtype, DE_RESIDUAL, a, s);
    if (rc) return rc;
  }
  GemvArgs a{};
  a.wt = m->whead_t; a.x = m->x; a.bias = m->bhead; a.out = m->logits; a.N = m->V; a.K = d;
  a.ln_scale = m->lnf_scale; a.ln_prev = nullptr; a.ln_shift = 0;                         // final LayerNorm fused (progen.py:220)
  int rc = gemv_dispatch(m->wdtype, DE_BIAS, a, s);
  if (rc) return rc;
  int threads = 32;
  while (threads < m->V) threads <<= 1;
  decode_sample_kernel<<<1, threads, m->V * sizeof(float), s>>>(m->logits, m->noise, m->seq, m->pos, m->logits_all, m->V, n,
                                                               m->top_k, do_sample);
  PG_LAUNCH_CHECK();
  return PROGEN_OK;
}

}  // extern "C"
