// KV-cached autoregressive decode as ONE persistent kernel (BASELINE config 5; reference utils.py:106-135, sample.py:66-71).
//
// Round 1 replayed a CUDA graph of ~150 tiny kernels per token: 1.0 ms / token for 103 MB of bf16 weights = 1.6 % of the
// HBM roofline, pure launch latency.  Here one cooperative kernel (one CTA per SM) generates every position of the
// launch: the phases of a layer (LN + shift + QKV + rotary + cache | windowed attention | out-proj + residual | LN + shift +
// FF-in + GLU/GELU | [gMLP: gate LN + causal spatial mix | SGU proj] | FF-out + residual) are separated by a grid barrier
// (one atomic + one acquire poll per CTA), the weights stream through every SM's warps with 16-byte loads, and the token
// loop, the sampler (top-k filter that keeps k-1 and zeroes the rest, Gumbel-max, `seq[pos+1] += id` — quirks Q5/Q6)
// and the position counter stay on the device: no host round trip, no launches.
//
// BATCH: `B` sequences advance in lock step ([B, 1] rows per step).  Every weight chunk a lane loads is used against all B
// activation rows (staged in shared memory), partial sums are reduced across lanes with a transposing butterfly (31
// shuffles per 32 values), so the step streams the weights ONCE for all sequences: decode throughput scales with B until
// the FMA pipe, not HBM, is the bound.  Sequence b samples position p+1 iff p+1 >= start[b] (its prime is kept before).
#include "common.cuh"
#include "../../include/progen_b200.h"

namespace {

constexpr int TPB = 256, WPB = TPB / 32;
constexpr int KC = 512;                 // activation columns staged per pass

// ------------------------------------------------------------------------------------------------ grid barrier
// monotonic counter: every CTA adds 1, then polls until all gridDim.x arrivals of this round are in
__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& round) {
  __syncthreads();
  if (threadIdx.x == 0) {
    ++round;
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned int target = round * gridDim.x;
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// sum over the 32 lanes of v[i], result for index i lands in lane i (v[0] of that lane)
template <int N> __device__ __forceinline__ void xreduce(float (&v)[32], int lane) {
  constexpr int H = N / 2;
  const bool up = (lane & H) != 0;
#pragma unroll
  for (int i = 0; i < H; ++i) {
    const float send = up ? v[i] : v[i + H];
    const float keep = up ? v[i + H] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, H);
  }
}
__device__ __forceinline__ float xreduce32(float (&v)[32], int lane) {
  xreduce<32>(v, lane); xreduce<16>(v, lane); xreduce<8>(v, lane); xreduce<4>(v, lane); xreduce<2>(v, lane);
  return v[0];
}

template <typename TW> __device__ __forceinline__ void load_w8(const TW* p, float (&w)[8]);
template <> __device__ __forceinline__ void load_w8<float>(const float* p, float (&w)[8]) { load_vec<8>(p, w); }
template <> __device__ __forceinline__ void load_w8<bf16>(const bf16* p, float (&w)[8]) { load_vec<8>(p, w); }

enum { EP_BIAS = 0, EP_ROTARY_CACHE = 1, EP_RESIDUAL = 2, EP_GLU = 3, EP_GELU = 4 };
enum { PRO_NONE = 0, PRO_LN = 1 };

struct Phase {
  const void* wt;          // [N(,x2 for GLU), K]
  const float* bias;       // [N] or null
  const float* xin;        // [B, ldx] input rows
  int ldx;
  float* out;              // [B, ldo]
  int ldo;
  int N, K;
  int epi;
  // prologue
  int pro;                 // PRO_LN: x <- shift(LN(x) * scale)
  const float* ln_scale;
  float* ln_prev;          // [B][2][K/2] token-shift state (read [pos&1], write [(pos+1)&1]); null: no shift
  // rotary / cache epilogue
  float* kcache; float* vcache; int inner, dim_head, n;
  const float* rot_sin; const float* rot_cos;
  int pos;
};

// One GEMV / skinny-GEMM phase over all B sequences.  BT = compile-time batch tile (B <= BT).
template <int BT, typename TW>
__device__ void gemv_phase(const Phase& ph, int B, float* xs /* smem [BT][KC] */, float* red /* smem scratch */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const TW* W = reinterpret_cast<const TW*>(ph.wt);
  const int npairs = ph.epi == EP_GLU ? ph.N : ph.N / 2;
  const int total_warps = gridDim.x * WPB;
  const int gw = blockIdx.x * WPB + warp;
  const int rounds = (npairs + total_warps - 1) / total_warps;
  const int nchunks = (ph.K + KC - 1) / KC;
  // LN statistics of every sequence's row (whole K), once per phase: warp b % WPB handles row b
  float* stat = red;                       // [BT][2]
  if (ph.pro == PRO_LN) {
    for (int b = warp; b < B; b += WPB) {
      const float* xr = ph.xin + (long long)b * ph.ldx;
      float s = 0.f;
      for (int k = lane * 4; k < ph.K; k += 128) { const float4 t = *reinterpret_cast<const float4*>(xr + k); s += (t.x + t.y) + (t.z + t.w); }
      s = warp_sum(s);
      const float mean = s / ph.K;
      float q = 0.f;
      for (int k = lane * 4; k < ph.K; k += 128) {
        const float4 t = *reinterpret_cast<const float4*>(xr + k);
        const float a0 = t.x - mean, a1 = t.y - mean, a2 = t.z - mean, a3 = t.w - mean;
        q += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
      }
      q = warp_sum(q);
      if (lane == 0) { stat[2 * b] = mean; stat[2 * b + 1] = rsqrtf(q / ph.K + 1e-5f); }
    }
    __syncthreads();
  }
  for (int rd = 0; rd < rounds; ++rd) {
    const int pair = rd * total_warps + gw;
    const bool active = pair < npairs;
    int r0 = 0, r1 = 0;
    if (active) {
      if (ph.epi == EP_GLU) { r0 = pair; r1 = pair + ph.N; }
      else { r0 = 2 * pair; r1 = r0 + 1; }
    }
    float acc0[BT], acc1[BT];
#pragma unroll
    for (int b = 0; b < BT; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
    for (int kc = 0; kc < nchunks; ++kc) {
      const int k0 = kc * KC, kn = min(KC, ph.K - k0);
      // ---- stage x[:, k0 .. k0+kn) (with the LN + shift prologue) into shared memory
      if (rd == 0 || nchunks > 1) {
        __syncthreads();
        const int half = ph.K >> 1;
        for (int idx = threadIdx.x; idx < B * (kn >> 2); idx += TPB) {
          const int b = idx / (kn >> 2), k = (idx % (kn >> 2)) * 4;
          float4 t = *reinterpret_cast<const float4*>(ph.xin + (long long)b * ph.ldx + k0 + k);
          if (ph.pro == PRO_LN) {
            const float mean = stat[2 * b], rstd = stat[2 * b + 1];
            const float4 sc = *reinterpret_cast<const float4*>(ph.ln_scale + k0 + k);
            t.x = (t.x - mean) * rstd * sc.x; t.y = (t.y - mean) * rstd * sc.y;
            t.z = (t.z - mean) * rstd * sc.z; t.w = (t.w - mean) * rstd * sc.w;
            if (ph.ln_prev && k0 + k < half) {
              float* st = ph.ln_prev + (long long)b * ph.K;            // [2][K/2]
              const float4 pv = *reinterpret_cast<const float4*>(st + (ph.pos & 1) * half + k0 + k);
              if (blockIdx.x == 0) *reinterpret_cast<float4*>(st + ((ph.pos + 1) & 1) * half + k0 + k) = t;
              t = pv;
            }
          }
          *reinterpret_cast<float4*>(xs + b * KC + k) = t;
        }
        __syncthreads();
      }
      if (!active) continue;
      // ---- this warp's two weight rows against every staged activation row
      const TW* w0 = W + (long long)r0 * ph.K + k0;
      const TW* w1 = W + (long long)r1 * ph.K + k0;
      for (int k = lane * 8; k < kn; k += 256) {
        float a[8], c[8];
        load_w8<TW>(w0 + k, a);
        load_w8<TW>(w1 + k, c);
#pragma unroll
        for (int b = 0; b < BT; ++b) {
          if (b < B) {
            const float4 x0 = *reinterpret_cast<const float4*>(xs + b * KC + k);
            const float4 x1 = *reinterpret_cast<const float4*>(xs + b * KC + k + 4);
            acc0[b] = fmaf(a[0], x0.x, fmaf(a[1], x0.y, fmaf(a[2], x0.z, fmaf(a[3], x0.w, acc0[b]))));
            acc0[b] = fmaf(a[4], x1.x, fmaf(a[5], x1.y, fmaf(a[6], x1.z, fmaf(a[7], x1.w, acc0[b]))));
            acc1[b] = fmaf(c[0], x0.x, fmaf(c[1], x0.y, fmaf(c[2], x0.z, fmaf(c[3], x0.w, acc1[b]))));
            acc1[b] = fmaf(c[4], x1.x, fmaf(c[5], x1.y, fmaf(c[6], x1.z, fmaf(c[7], x1.w, acc1[b]))));
          }
        }
      }
    }
    if (!active) continue;
    // ---- reduce over lanes; lane L ends with the sums of sequence (g * 32 + L), g = 0 .. BT/32-1  (BT < 32: all lanes hold b)
    constexpr int NG = BT >= 32 ? BT / 32 : 1;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float s0, s1;
      int b;
      if constexpr (BT >= 32) {
        float v0[32], v1[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { v0[i] = acc0[g * 32 + i]; v1[i] = acc1[g * 32 + i]; }
        s0 = xreduce32(v0, lane);
        s1 = xreduce32(v1, lane);
        b = g * 32 + lane;
      } else {
        s0 = 0.f; s1 = 0.f; b = lane;
#pragma unroll
        for (int i = 0; i < BT; ++i) {
          const float t0 = warp_sum(acc0[i]), t1 = warp_sum(acc1[i]);
          if (lane == i) { s0 = t0; s1 = t1; }
        }
      }
      if (b >= B) continue;
      if (ph.bias) { s0 += ph.bias[r0]; s1 += ph.bias[r1]; }
      float* o = ph.out + (long long)b * ph.ldo;
      if (ph.epi == EP_BIAS) { o[r0] = s0; o[r1] = s1; }
      else if (ph.epi == EP_RESIDUAL) { o[r0] += s0; o[r1] += s1; }
      else if (ph.epi == EP_GELU) { o[r0] = gelu_tanh(s0); o[r1] = gelu_tanh(s1); }
      else if (ph.epi == EP_GLU) { o[r0] = s0 * gelu_tanh(s1); }
      else {  // EP_ROTARY_CACHE: rotary on q, k AND v (progen.py:87); k, v rows go to the caches at position pos
        const int hd = ph.dim_head >> 1, j = (r0 % ph.dim_head) >> 1;
        const float sn = ph.rot_sin[ph.pos * hd + j], cs = ph.rot_cos[ph.pos * hd + j];
        const float o0 = s0 * cs - s1 * sn, o1 = s1 * cs + s0 * sn;
        const int sec = r0 / ph.inner, c = r0 % ph.inner;
        float* dst = sec == 0 ? o + c
                              : (sec == 1 ? ph.kcache : ph.vcache) + ((long long)b * ph.n + ph.pos) * ph.inner + c;
        dst[0] = o0; dst[1] = o1;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention
// task = (sequence, head, slice of 32 keys): partial (m, l, o[dh]) -> att_part; the warp that finishes a (sequence, head)'s
// last slice merges the partials (plus window 0's w zero keys with logit 0, quirk Q1) into att[b, head * dh ..].
__device__ void attention_phase(const progen_decode_run_t& r, const progen_decode_layer_t& L, int pos, float* sq /* smem [WPB][dh] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int dh = r.dim_head, w = r.window, I = r.inner;
  const int win = pos / w, i = pos % w;
  const int key0 = win > 0 ? (win - 1) * w : 0;
  const int nreal = (win > 0 ? w : 0) + i + 1;
  const int nsl = (nreal + 31) / 32;
  const int KS = (2 * w + 31) / 32;                         // slots per (b, head) in att_part
  const int tasks = r.B * r.heads * nsl;
  const float scale = rsqrtf((float)dh);
  float* q_s = sq + warp * dh;
  for (int t = blockIdx.x * WPB + warp; t < tasks; t += gridDim.x * WPB) {
    const int sl = t % nsl, bh = t / nsl, hh = bh % r.heads, b = bh / r.heads;
    const float* qv = r.q + (long long)b * I + hh * dh;
    for (int c = lane; c < dh; c += 32) q_s[c] = qv[c];
    __syncwarp();
    const int j = sl * 32 + lane;
    const bool valid = j < nreal;
    const float* kr = L.kcache + ((long long)b * r.n + key0 + (valid ? j : 0)) * I + hh * dh;
    float s = 0.f;
    for (int c = 0; c < dh; c += 4) {
      const float4 kv = *reinterpret_cast<const float4*>(kr + c);
      s = fmaf(kv.x, q_s[c], s); s = fmaf(kv.y, q_s[c + 1], s); s = fmaf(kv.z, q_s[c + 2], s); s = fmaf(kv.w, q_s[c + 3], s);
    }
    s = valid ? s * scale : -INFINITY;
    const float m = warp_max(s);
    const float p = valid ? expf(s - m) : 0.f;
    const float l = warp_sum(p);
    // o[c] = sum_j p_j v_j[c]: lanes own channels (c = lane, lane + 32, ...), p_j broadcast key by key
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
    const int nk = min(32, nreal - sl * 32);
    const float* vb = L.vcache + ((long long)b * r.n + key0 + sl * 32) * I + hh * dh;
    for (int jj = 0; jj < nk; ++jj) {
      const float pj = __shfl_sync(0xffffffffu, p, jj);
      const float* vr = vb + (long long)jj * I;
      if (lane < dh) o0 = fmaf(pj, vr[lane], o0);
      if (lane + 32 < dh) o1 = fmaf(pj, vr[lane + 32], o1);
      if (lane + 64 < dh) o2 = fmaf(pj, vr[lane + 64], o2);
      if (lane + 96 < dh) o3 = fmaf(pj, vr[lane + 96], o3);
    }
    float* part = r.att_part + ((long long)bh * KS + sl) * (dh + 2);
    if (lane < dh) part[2 + lane] = o0;
    if (lane + 32 < dh) part[2 + lane + 32] = o1;
    if (lane + 64 < dh) part[2 + lane + 64] = o2;
    if (lane + 96 < dh) part[2 + lane + 96] = o3;
    if (lane == 0) { part[0] = m; part[1] = l; }
    // last slice of this (b, head) to finish merges
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atomicAdd(r.att_count + bh, 1) == nsl - 1;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      __threadfence();
      const float* pb = r.att_part + (long long)bh * KS * (dh + 2);
      float M = win == 0 ? 0.f : -INFINITY;               // zero look-back keys of window 0: logit 0 (quirk Q1)
      for (int k = 0; k < nsl; ++k) M = fmaxf(M, __ldcg(pb + k * (dh + 2)));
      float Lt = win == 0 ? (float)w * expf(-M) : 0.f;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int k = 0; k < nsl; ++k) {
        const float* pk = pb + k * (dh + 2);
        const float f = expf(__ldcg(pk) - M);
        Lt += __ldcg(pk + 1) * f;
        if (lane < dh) a0 = fmaf(f, __ldcg(pk + 2 + lane), a0);
        if (lane + 32 < dh) a1 = fmaf(f, __ldcg(pk + 2 + lane + 32), a1);
        if (lane + 64 < dh) a2 = fmaf(f, __ldcg(pk + 2 + lane + 64), a2);
        if (lane + 96 < dh) a3 = fmaf(f, __ldcg(pk + 2 + lane + 96), a3);
      }
      const float inv = 1.f / Lt;
      float* ao = r.att + (long long)b * I + hh * dh;
      if (lane < dh) ao[lane] = a0 * inv;
      if (lane + 32 < dh) ao[lane + 32] = a1 * inv;
      if (lane + 64 < dh) ao[lane + 64] = a2 * inv;
      if (lane + 96 < dh) ao[lane + 96] = a3 * inv;
      if (lane == 0) r.att_count[bh] = 0;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ SGU (progen.py:166-184)
// a = gelu(proj_in) = [xs | gate] (C channels each).  gn = LN(gate) * scale -> history[b][pos]; gate' = sum_{k<=pos} W[pos,k]
// history[b][k] + bias[pos]; sg = xs * gate'.  Task = (sequence, block of 128 channels); warps split the k range.
__device__ void sgu_phase(const progen_decode_run_t& r, const progen_decode_layer_t& L, int pos, float* red /* smem [WPB][128] + stats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = r.hid / 2, n = r.n;
  const int cblocks = C / 128;
  const int tasks = r.B * cblocks;
  float* stat = red + WPB * 128;
  for (int t = blockIdx.x; t < tasks; t += gridDim.x) {
    const int cb = t % cblocks, b = t / cblocks;
    const float* gate = r.u + (long long)b * r.hid + C;
    // LN statistics of the gate row (every task recomputes them: C floats)
    __syncthreads();
    {
      float s = 0.f;
      for (int c = threadIdx.x; c < C; c += TPB) s += gate[c];
      s = warp_sum(s);
      if (lane == 0) red[warp] = s;
      __syncthreads();
      if (threadIdx.x == 0) { float tt = 0.f; for (int k = 0; k < WPB; ++k) tt += red[k]; stat[0] = tt / C; }
      __syncthreads();
      const float mean = stat[0];
      float qq = 0.f;
      for (int c = threadIdx.x; c < C; c += TPB) { const float u = gate[c] - mean; qq += u * u; }
      qq = warp_sum(qq);
      if (lane == 0) red[warp] = qq;
      __syncthreads();
      if (threadIdx.x == 0) { float tt = 0.f; for (int k = 0; k < WPB; ++k) tt += red[k]; stat[1] = rsqrtf(tt / C + 1e-5f); }
      __syncthreads();
    }
    const float mean = stat[0], rstd = stat[1];
    const int c0 = cb * 128 + lane * 4;
    float4 gnow;
    {
      const float4 gv = *reinterpret_cast<const float4*>(gate + c0);
      const float4 sc = *reinterpret_cast<const float4*>(L.sgu_ln_scale + c0);
      gnow.x = (gv.x - mean) * rstd * sc.x; gnow.y = (gv.y - mean) * rstd * sc.y;
      gnow.z = (gv.z - mean) * rstd * sc.z; gnow.w = (gv.w - mean) * rstd * sc.w;
    }
    float* hist = L.gn_hist + (long long)b * n * C;
    if (warp == 0) *reinterpret_cast<float4*>(hist + (long long)pos * C + c0) = gnow;
    const float* wrow = L.sgu_w + (long long)pos * n;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = warp; k < pos; k += WPB) {                       // earlier positions from the history
      const float wk = __ldg(wrow + k);
      const float4 h = *reinterpret_cast<const float4*>(hist + (long long)k * C + c0);
      acc.x = fmaf(wk, h.x, acc.x); acc.y = fmaf(wk, h.y, acc.y); acc.z = fmaf(wk, h.z, acc.z); acc.w = fmaf(wk, h.w, acc.w);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(red + warp * 128 + lane * 4) = acc;
    __syncthreads();
    if (warp == 0) {
      float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < WPB; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(red + k * 128 + lane * 4);
        tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
      }
      const float wp = __ldg(wrow + pos), bp = L.sgu_b[pos];    // the current position's own term + spatial bias
      const float4 xv = *reinterpret_cast<const float4*>(r.u + (long long)b * r.hid + c0);
      float4 o;
      o.x = xv.x * (fmaf(wp, gnow.x, tot.x) + bp); o.y = xv.y * (fmaf(wp, gnow.y, tot.y) + bp);
      o.z = xv.z * (fmaf(wp, gnow.z, tot.z) + bp); o.w = xv.w * (fmaf(wp, gnow.w, tot.w) + bp);
      *reinterpret_cast<float4*>(r.sg + (long long)b * C + c0) = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------ sampler (one sequence per CTA)
// utils.py:97-129: top-k filter keeps logits > (k-th largest), the rest become 0.0 and lose their noise; argmax(logits +
// gumbel) (first maximal index); seq[pos + 1] += index (ADD, quirk Q5).
__device__ void sample_phase(const progen_decode_run_t& r, int pos, float* sv /* smem [V] */, float* red) {
  const int t = threadIdx.x, V = r.V;
  for (int b = blockIdx.x; b < r.B; b += gridDim.x) {
    __syncthreads();
    const float* lg = r.logits + (long long)b * V;
    if (r.logits_all) for (int c = t; c < V; c += TPB) r.logits_all[((long long)b * r.n + pos) * V + c] = lg[c];
    if (pos + 1 >= r.n || pos + 1 < r.start[b]) continue;                      // the prime is kept; nothing after the end
    for (int c = t; c < V; c += TPB) sv[c] = lg[c];
    __syncthreads();
    float kth = -INFINITY;
    if (r.top_k > 0) {
      for (int c = t; c < V; c += TPB) {
        const float v = sv[c];
        int gt = 0, ge = 0;
        for (int j = 0; j < V; ++j) { gt += sv[j] > v; ge += sv[j] >= v; }
        if (gt < r.top_k && r.top_k <= ge) red[0] = v;                         // the k-th largest value (with multiplicity)
      }
      __syncthreads();
      kth = red[0];
    }
    __syncthreads();
    for (int c = t; c < V; c += TPB) {
      const float v = sv[c];
      const bool keep = r.top_k > 0 ? v > kth : true;
      const float nz = r.noise ? r.noise[((long long)b * r.n + pos) * V + c] : 0.f;
      sv[c] = keep ? v + nz : 0.f;
    }
    __syncthreads();
    if (t == 0) {
      int best = 0;
      float bv = sv[0];
      for (int j = 1; j < V; ++j) if (sv[j] > bv) { bv = sv[j]; best = j; }
      r.seq[(long long)b * r.n + pos + 1] += best;
    }
  }
}

template <int BT, typename TW>
__global__ void __launch_bounds__(TPB, 1) decode_persistent_kernel(const progen_decode_run_t r) {
  extern __shared__ float smem[];
  float* xs = smem;                                    // [BT][KC]
  float* red = smem + BT * KC;                         // scratch: WPB*128 + 2*BT + 16
  unsigned int round = 0;
  const int d = r.d, I = r.inner, hid = r.hid, B = r.B;
  for (int step = 0; step < r.nsteps; ++step) {
    const int pos = r.pos0 + step;
    // ---- embedding: x[b] = embed[clamp(seq[b][pos])]
    for (int idx = blockIdx.x * TPB + threadIdx.x; idx < B * (d >> 2); idx += gridDim.x * TPB) {
      const int b = idx / (d >> 2), c = (idx % (d >> 2)) * 4;
      int id = r.seq[(long long)b * r.n + pos];
      id = id < 0 ? 0 : (id >= r.V ? r.V - 1 : id);
      *reinterpret_cast<float4*>(r.x + (long long)b * d + c) = *reinterpret_cast<const float4*>(r.embed + (long long)id * d + c);
    }
    grid_sync(r.grid_bar, round);
    for (int li = 0; li < r.depth; ++li) {
      const progen_decode_layer_t& L = r.layers[li];
      Phase ph{};
      // ---- LN + shift + QKV + rotary + cache
      ph.wt = L.wqkv_t; ph.bias = nullptr; ph.xin = r.x; ph.ldx = d; ph.out = r.q; ph.ldo = I; ph.N = 3 * I; ph.K = d; ph.epi = EP_ROTARY_CACHE;
      ph.pro = PRO_LN; ph.ln_scale = L.ln1_scale; ph.ln_prev = r.shift_tokens ? L.shift1 : nullptr;
      ph.kcache = L.kcache; ph.vcache = L.vcache; ph.inner = I; ph.dim_head = r.dim_head; ph.n = r.n; ph.rot_sin = r.rot_sin; ph.rot_cos = r.rot_cos; ph.pos = pos;
      gemv_phase<BT, TW>(ph, B, xs, red);
      grid_sync(r.grid_bar, round);
      attention_phase(r, L, pos, red);
      grid_sync(r.grid_bar, round);
      // ---- out-proj + residual
      ph = Phase{};
      ph.wt = L.wo_t; ph.bias = L.bo; ph.xin = r.att; ph.ldx = I; ph.out = r.x; ph.ldo = d; ph.N = d; ph.K = I; ph.epi = EP_RESIDUAL; ph.pos = pos;
      gemv_phase<BT, TW>(ph, B, xs, red);
      grid_sync(r.grid_bar, round);
      // ---- LN + shift + FF-in (+ GLU / GELU)
      ph = Phase{};
      ph.wt = L.win_t; ph.bias = L.bin; ph.xin = r.x; ph.ldx = d; ph.out = r.u; ph.ldo = hid; ph.N = hid; ph.K = d;
      ph.epi = L.kind == 0 ? EP_GLU : EP_GELU; ph.pro = PRO_LN; ph.ln_scale = L.ln2_scale; ph.ln_prev = r.shift_tokens ? L.shift2 : nullptr; ph.pos = pos;
      gemv_phase<BT, TW>(ph, B, xs, red);
      grid_sync(r.grid_bar, round);
      const float* last = r.u;
      int last_k = hid, last_ld = hid;
      if (L.kind == 2) {
        sgu_phase(r, L, pos, red);
        grid_sync(r.grid_bar, round);
        ph = Phase{};
        ph.wt = L.sgu_proj_t; ph.bias = L.sgu_proj_b; ph.xin = r.sg; ph.ldx = hid / 2; ph.out = r.pj; ph.ldo = hid / 2; ph.N = hid / 2; ph.K = hid / 2;
        ph.epi = EP_BIAS; ph.pos = pos;
        gemv_phase<BT, TW>(ph, B, xs, red);
        grid_sync(r.grid_bar, round);
        last = r.pj; last_k = hid / 2; last_ld = hid / 2;
      }
      // ---- FF-out + residual
      ph = Phase{};
      ph.wt = L.wout_t; ph.bias = L.bout; ph.xin = last; ph.ldx = last_ld; ph.out = r.x; ph.ldo = d; ph.N = d; ph.K = last_k; ph.epi = EP_RESIDUAL; ph.pos = pos;
      gemv_phase<BT, TW>(ph, B, xs, red);
      grid_sync(r.grid_bar, round);
    }
    // ---- final LN + logits (progen.py:219-222), then the sampler
    Phase ph{};
    ph.wt = r.whead_t; ph.bias = r.bhead; ph.xin = r.x; ph.ldx = d; ph.out = r.logits; ph.ldo = r.V; ph.N = r.V; ph.K = d; ph.epi = EP_BIAS;
    ph.pro = PRO_LN; ph.ln_scale = r.lnf_scale; ph.ln_prev = nullptr; ph.pos = pos;
    gemv_phase<BT, TW>(ph, B, xs, red);
    grid_sync(r.grid_bar, round);
    sample_phase(r, pos, xs, red);
    grid_sync(r.grid_bar, round);
  }
}

template <int BT, typename TW>
int launch_run(const progen_decode_run_t& r, cudaStream_t s) {
  const size_t smem = (size_t)(BT * KC + WPB * 128 + 2 * BT + 64) * sizeof(float);
  auto kern = decode_persistent_kernel<BT, TW>;
  static bool once = false;
  if (!once) {
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once = true;
  }
  int per_sm = 0;
  PG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TPB, smem));
  PG_CHECK_ARG(per_sm >= 1);
  const int grid = pg_num_sms();
  void* args[] = {(void*)&r};
  PG_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(TPB), args, smem, s));
  __atomic_fetch_add(&g_progen_launches, 1ull, __ATOMIC_RELAXED);
  return PROGEN_OK;
}

}  // namespace

extern "C" {

// Consume positions pos0 .. pos0 + nsteps - 1 of all B sequences in ONE kernel.  `grid_bar` and `att_count` must be zero on entry
// (the kernel leaves att_count zero; the caller re-zeroes grid_bar before the next launch).
int progen_decode_run(const progen_decode_run_t* r, void* stream) {
  PG_CHECK_ARG(r != nullptr && r->layers != nullptr && r->depth > 0 && r->B >= 1 && r->B <= 64 && r->nsteps >= 0);
  PG_CHECK_ARG(r->d % 8 == 0 && r->inner % 8 == 0 && r->hid % 256 == 0 && r->V % 2 == 0 && r->V <= KC);
  PG_CHECK_ARG(r->dim_head % 4 == 0 && r->dim_head <= 128 && r->pos0 >= 0 && r->pos0 + r->nsteps <= r->n);
  PG_CHECK_ARG(r->grid_bar != nullptr && r->att_count != nullptr && r->att_part != nullptr);
  if (r->nsteps == 0) return PROGEN_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool bf = r->wdtype == PG_BF16;
  if (r->B == 1) return bf ? launch_run<1, bf16>(*r, s) : launch_run<1, float>(*r, s);
  if (r->B <= 8) return bf ? launch_run<8, bf16>(*r, s) : launch_run<8, float>(*r, s);
  if (r->B <= 32) return bf ? launch_run<32, bf16>(*r, s) : launch_run<32, float>(*r, s);
  return bf ? launch_run<64, bf16>(*r, s) : launch_run<64, float>(*r, s);
}

}  // extern "C"
