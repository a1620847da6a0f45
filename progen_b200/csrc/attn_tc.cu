// Sliding-window attention FORWARD on tcgen05 tensor cores (sm_100a): reference progen.py:88-102, bf16, dim_head 64,
// window % 128 == 0.  This is the kernel BASELINE's north_star describes: K/V tiles staged by TMA into 128B-swizzled
// shared memory, QK^T and PV as tcgen05.mma with accumulators in TMEM, softmax by threads that own one query row each
// (TMEM lane == row, so row max / row sum need no shuffles).
//
//   warp 0     : TMA producer  — Q tile once per work item, K|V tiles through a 3-stage ring
//   warp 1     : MMA issuer    — S_j = Q K_j^T (128x128x64) into one of two TMEM S buffers; O_j = P_j V_j (128x64x128) into
//                                one of two TMEM O buffers; QK of tile j+1 is issued before PV of tile j so the tensor pipe
//                                works while the softmax warps process tile j
//   warp 2     : TMEM allocator
//   warps 4..11: softmax       — two passes over S_j straight from TMEM (max, then exp2), P_j written as bf16 into the
//                                K-major swizzled smem layout the PV MMA reads; running (m, l) and the output row O live in
//                                registers: O = O * exp2(m_old - m_new) + (P_j V_j read back from TMEM), so TMEM is never
//                                rescaled in place
// Persistent over (batch, head, 128-query tile) work items.  Window 0's zero look-back keys (reference quirk Q1) enter
// analytically: m starts at 0 and l at w.
#include "tc_ptx.cuh"
#include "../../include/progen_b200.h"

// attn_tc_pair.cu
int attn_fwd_pair_launch(const void* qkv, void* out, float* lse, int B, int seq_len, int window, int heads, cudaStream_t stream);
int attn_fwd_ts_launch(const void* qkv, void* out, float* lse, int B, int seq_len, int window, int heads, cudaStream_t stream);

namespace {

using namespace tc;

constexpr int BQ = 128, BKV = 128, DH = 64;
constexpr int KV_STAGES = 3;
constexpr int Q_BYTES = BQ * DH * 2;              // 16 KiB
constexpr int K_BYTES = BKV * DH * 2;             // 16 KiB
constexpr int KV_BYTES = 2 * K_BYTES;             // K then V
constexpr int P_BYTES = BQ * BKV * 2;             // 32 KiB: two [128 x 64] K-major sub-tiles
constexpr int BAR_BYTES = 144 + 2 * 2 * BQ * 4 + 2 * BQ * 4 + 112;   // barriers + TMEM slot (144 B), row-max exchange [2][2][128], row-sum exchange [2][128]
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * KV_BYTES + 2 * P_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 512;                    // S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnDev {
  int B, n, w, h;
  bf16* out;
  float* lse;
};

struct Item { int b, hh, q0, win, nprev, ntiles; };

__device__ __forceinline__ bool decode_item(const AttnDev& a, int wi, Item& it) {
  const int qtiles = a.n / BQ;
  if (wi >= a.B * a.h * qtiles) return false;
  const int qt = wi % qtiles;
  const int r = wi / qtiles;
  it.hh = r % a.h;
  it.b = r / a.h;
  it.q0 = qt * BQ;
  it.win = it.q0 / a.w;
  const int i0 = it.q0 % a.w;
  it.nprev = it.win > 0 ? a.w / BKV : 0;
  it.ntiles = it.nprev + i0 / BKV + 1;                     // the last tile is the causal diagonal tile
  return true;
}
__device__ __forceinline__ int key_pos(const AttnDev& a, const Item& it, int kt) {
  return kt < it.nprev ? (it.win - 1) * a.w + kt * BKV : it.win * a.w + (kt - it.nprev) * BKV;
}

__global__ void __launch_bounds__(384, 1) attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnDev a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + Q_BYTES;
  const uint32_t sP = sKV + KV_STAGES * KV_BYTES;
  const uint32_t bars = sP + 2 * P_BYTES;
  // barriers (8 B each)
  const uint32_t q_full = bars, q_empty = bars + 8;
  auto kv_full = [&](int s) { return bars + 16 + 8 * s; };
  auto kv_empty = [&](int s) { return bars + 16 + 8 * (KV_STAGES + s); };
  auto s_full = [&](int i) { return bars + 64 + 8 * i; };
  auto s_empty = [&](int i) { return bars + 80 + 8 * i; };
  auto p_full = [&](int i) { return bars + 96 + 8 * i; };
  auto o_full = [&](int i) { return bars + 112 + 8 * i; };
  const uint32_t tmem_slot = bars + 128;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));          // generic pointer to the aligned base

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = a.h * DH;

  if (warp == 0 && lane == 0) prefetch_tensormap(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(s_full(i), 1);
      mbar_init(s_empty(i), 8);      // one arrival per softmax warp
      mbar_init(p_full(i), 8);
      mbar_init(o_full(i), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  if (warp == 0) {
    // ============================================================================ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t kv_phase = 0, q_phase = 0;
      Item it;
      for (int wi = blockIdx.x; decode_item(a, wi, it); wi += gridDim.x) {
        const int row0 = it.b * a.n;
        mbar_wait(q_empty, q_phase ^ 1);
        mbar_expect_tx(q_full, Q_BYTES);
        tma_load_2d(sQ, &tmap_qkv, q_full, it.hh * DH, row0 + it.q0);
        q_phase ^= 1;
        for (int kt = 0; kt < it.ntiles; ++kt) {
          mbar_wait(kv_empty(stage), kv_phase ^ 1);
          const uint32_t dst = sKV + stage * KV_BYTES;
          const int kp = row0 + key_pos(a, it, kt);
          mbar_expect_tx(kv_full(stage), KV_BYTES);
          tma_load_2d(dst, &tmap_qkv, kv_full(stage), I + it.hh * DH, kp);
          tma_load_2d(dst + K_BYTES, &tmap_qkv, kv_full(stage), 2 * I + it.hh * DH, kp);
          if (++stage == KV_STAGES) { stage = 0; kv_phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc(BQ, BKV, false, false);     // S[128 x 128] = Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = make_idesc(BQ, DH, false, true);       // O[128 x 64]  = P (K-major) x V (MN-major)
      int stage = 0;
      uint32_t kv_phase = 0, q_phase = 0;
      uint32_t g = 0;                                                      // global tile counter -> S/P/O buffer and parity
      Item it;
      auto issue_qk = [&](int st, uint32_t gi) {
        const uint32_t buf = gi & 1;
        if (gi >= 2) mbar_wait(s_empty(buf), ((gi - 2) >> 1) & 1);         // softmax finished reading S of tile gi-2
        tcgen05_fence_after();
        const uint64_t ad = make_smem_desc<false>(sQ);
        const uint64_t bd = make_smem_desc<false>(sKV + st * KV_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + buf * BKV, ad + 2 * k, bd + 2 * k, idesc_qk, k > 0);
        tcgen05_commit(s_full(buf));
      };
      for (int wi = blockIdx.x; decode_item(a, wi, it); wi += gridDim.x) {
        mbar_wait(q_full, q_phase);
        q_phase ^= 1;
        int qk_stage = stage;
        uint32_t qk_phase = kv_phase;
        // S_0
        mbar_wait(kv_full(qk_stage), qk_phase);
        issue_qk(qk_stage, g);
        if (++qk_stage == KV_STAGES) { qk_stage = 0; qk_phase ^= 1; }
        for (int j = 0; j < it.ntiles; ++j) {
          if (j + 1 < it.ntiles) {                                         // S_{j+1} overlaps softmax of tile j
            mbar_wait(kv_full(qk_stage), qk_phase);
            issue_qk(qk_stage, g + j + 1);
            if (++qk_stage == KV_STAGES) { qk_stage = 0; qk_phase ^= 1; }
          } else {
            tcgen05_commit(q_empty);                                       // every QK of this item has been issued
          }
          const uint32_t gj = g + j, buf = gj & 1;
          mbar_wait(p_full(buf), (gj >> 1) & 1);                           // P_j is in shared memory
          tcgen05_fence_after();
          const uint32_t pbase = sP + buf * P_BYTES;
          const uint64_t vd = make_smem_desc<true>(sKV + stage * KV_BYTES + K_BYTES);
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {
            const uint64_t pd = make_smem_desc<false>(pbase + (k >> 2) * (BQ * 128)) + 2 * (k & 3);
            umma_bf16(tmem_base + 256 + buf * DH, pd, vd + (uint64_t)(k * (2048 >> 4)), idesc_pv, k > 0);
          }
          tcgen05_commit(o_full(buf));
          tcgen05_commit(kv_empty(stage));                                 // K_j, V_j free once these MMAs retire
          if (++stage == KV_STAGES) { stage = 0; kv_phase ^= 1; }
        }
        g += it.ntiles;
      }
    }
  } else if (warp >= 4) {
    // ============================================================================ softmax: 8 warps, thread == (query row, half)
    // warps 4..7 take key columns [0, 64) of the S tile and output channels [0, 32); warps 8..11 the other halves.  The two
    // threads of a row only exchange their partial row maximum (shared memory + a 256-thread named barrier); the partial
    // row sums are combined once at the end.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float sc = 0.125f * LOG2E;                                        // 1/sqrt(64) in log2 units
    float* xmax = reinterpret_cast<float*>(gen_base + (bars - smem_base) + 144);   // [2 tiles][2 halves][128 rows]
    uint32_t g = 0;
    Item it;
    for (int wi = blockIdx.x; decode_item(a, wi, it); wi += gridDim.x) {
      float m_run = it.win == 0 ? 0.f : -INFINITY;                          // quirk Q1: w zero keys with logit 0
      float l_run = (it.win == 0 && half == 0) ? (float)a.w : 0.f;          // partial sum; halves are added at the end
      float o[DH / 2];
#pragma unroll
      for (int i = 0; i < DH / 2; ++i) o[i] = 0.f;
      float corr_pending = 1.f;
      for (int j = 0; j <= it.ntiles; ++j) {
        if (j < it.ntiles) {
          const uint32_t gj = g + j, buf = gj & 1;
          const bool diag = j == it.ntiles - 1;
          mbar_wait(s_full(buf), (gj >> 1) & 1);
          tcgen05_fence_after();
          const uint32_t s_addr = tmem_base + buf * BKV + half * 64 + lane_addr;
          float v[64];
          float mx = -INFINITY;
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int c = half * 2 + cc;                                    // 32-key chunk index inside the tile
            if (diag && c > q) {                                            // warp-uniform: above the diagonal for all rows
#pragma unroll
              for (int i = 0; i < 32; ++i) v[cc * 32 + i] = -INFINITY;
            } else {
              float t[32];
              tmem_ld32(s_addr + cc * 32, t);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = (diag && c * 32 + i > row) ? -INFINITY : t[i];
                v[cc * 32 + i] = x;
                mx = fmaxf(mx, x);
              }
            }
          }
          // S_j is in registers: release the TMEM buffer, then exchange the row maximum with the other half
          tcgen05_fence_before();
          xmax[(buf * 2 + half) * BQ + row] = mx;
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (lane == 0) mbar_arrive(s_empty(buf));
          mx = fmaxf(mx, xmax[(buf * 2 + (half ^ 1)) * BQ + row]);
          const float m_new = fmaxf(m_run, mx * sc);
          const float corr = ex2_approx(m_run - m_new);                     // ex2(-inf) = 0 on the first tile
          float rsum = 0.f;
          uint8_t* prow = gen_base + (sP - smem_base) + buf * P_BYTES + half * (BQ * 128) + row * 128;   // sub-tile == half
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float p[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              p[i] = ex2_approx(v[ch * 8 + i] * sc - m_new);                // masked entries: ex2(-inf) = 0
              rsum += p[i];
            }
            uint4 t;
            t.x = pack_bf16x2(p[0], p[1]); t.y = pack_bf16x2(p[2], p[3]); t.z = pack_bf16x2(p[4], p[5]); t.w = pack_bf16x2(p[6], p[7]);
            *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = t;
          }
          l_run = l_run * corr + rsum;
          m_run = m_new;
          fence_proxy_async();                                              // generic-proxy smem writes -> async proxy (MMA)
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full(buf));
          if (j > 0) {                                                      // consume P_{j-1} V_{j-1} (this thread's 32 channels)
            const uint32_t gp = gj - 1, pb = gp & 1;
            mbar_wait(o_full(pb), (gp >> 1) & 1);
            tcgen05_fence_after();
            float t[32];
            tmem_ld32(tmem_base + 256 + pb * DH + half * 32 + lane_addr, t);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_pending + t[i];
          }
          corr_pending = corr;
        } else {
          const uint32_t gp = g + it.ntiles - 1, pb = gp & 1;
          mbar_wait(o_full(pb), (gp >> 1) & 1);
          tcgen05_fence_after();
          float t[32];
          tmem_ld32(tmem_base + 256 + pb * DH + half * 32 + lane_addr, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = o[i] * corr_pending + t[i];
        }
      }
      g += it.ntiles;
      // combine the two partial row sums, then O / l -> bf16 (32 channels per thread); lse in natural-log units
      float* xsum = xmax + 2 * 2 * BQ;
      xsum[half * BQ + row] = l_run;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float l_tot = l_run + xsum[(half ^ 1) * BQ + row];
      const long long t = (long long)it.b * a.n + it.q0 + row;
      const float inv = 1.f / l_tot;
      bf16* op = a.out + t * I + it.hh * DH + half * 32;
#pragma unroll
      for (int c = 0; c < DH / 2; c += 8) {
        uint4 u;
        u.x = pack_bf16x2(o[c] * inv, o[c + 1] * inv); u.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
        u.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv); u.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
        *reinterpret_cast<uint4*>(op + c) = u;
      }
      if (half == 0) a.lse[t * a.h + it.hh] = m_run * LN2 + logf(l_tot);
      tcgen05_fence_before();                                               // order the TMEM reads before the next item's MMAs
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace

extern "C" {

// tcgen05 forward; same contract as progen_local_attn_fwd but requires window % 128 == 0.
int progen_local_attn_fwd_tc(const void* qkv, void* out, float* lse, int B, int seq_len, int window, int heads, int dim_head,
                             void* stream) {
  PG_CHECK_ARG(B > 0 && heads > 0 && dim_head == DH && window % 128 == 0 && seq_len % window == 0);
  PG_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15) == 0);
  // round-2 kernel (P and O in tensor memory, attn_fwd_ts.cu) when the window holds whole pairs of query tiles
  const int rc_ts = attn_fwd_ts_launch(qkv, out, lse, B, seq_len, window, heads, (cudaStream_t)stream);
  if (rc_ts <= 0) return rc_ts;
  // round-1 form: two query tiles per CTA with independent softmax groups (attn_tc_pair.cu)
  const int rc_pair = attn_fwd_pair_launch(qkv, out, lse, B, seq_len, window, heads, (cudaStream_t)stream);
  if (rc_pair <= 0) return rc_pair;
  const long long T = (long long)B * seq_len;
  const int I = heads * DH;
  CUtensorMap tm;
  int rc = pg_tensor_map_2d_bf16(qkv, 3ull * I, (uint64_t)T, 3ull * I, DH, BQ, &tm);
  if (rc) return rc;
  static bool once = false;
  if (!once) {
    PG_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    once = true;
  }
  AttnDev a{B, seq_len, window, heads, (bf16*)out, lse};
  const long long items = (long long)B * heads * (seq_len / BQ);
  const int grid = (int)(items < pg_num_sms() ? items : pg_num_sms());
  attn_fwd_tc_kernel<<<grid, 384, SMEM_BYTES, (cudaStream_t)stream>>>(tm, a);
  PG_LAUNCH_CHECK();
  return PROGEN_OK;
}

}  // extern "C"
