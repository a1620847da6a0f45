// Sliding-window attention FORWARD, tcgen05, two query tiles per CTA (window % 256 == 0): reference progen.py:88-102.
//
// attn_tc.cu processes one 128-query tile at a time with its 8 softmax warps in lockstep (thread = (row, half of the
// keys); the halves exchange the row maximum through shared memory and a 256-thread barrier every tile).  ncu shows that
// kernel latency-bound: every K/V tile is a serial chain  MMA -> TMEM read -> max -> exchange -> exp2 -> P to smem ->
// MMA  with nothing else to issue meanwhile.  Here a CTA owns the PAIR of adjacent query tiles (A = rows q0..q0+127,
// B = q0+128..q0+255 of one window), and each gets its own group of 4 softmax warps (thread == query row, TMEM lane ==
// row, so row max / row sum need no communication at all):
//
//   warp 0     : TMA producer — Q_A, Q_B once per pair; K|V tiles through a 3-stage ring, shared by both groups
//                (the two tiles see the same look-back window and the same own-window tiles: half the K/V traffic)
//   warp 1     : MMA issuer   — S_g = Q_g K_j^T (128x128x64) into group g's TMEM S buffer, O_g = P_g V_j (128x64x128);
//                PV_g(j) and S_g(j+1) are issued back to back as soon as group g has published P_g(j), so the tensor
//                pipe serves one group while the other group is in its softmax
//   warp 2     : TMEM allocator
//   warps 4..7 : softmax group A, warps 8..11: group B — per tile two passes over S straight from TMEM (row max, then
//                exp2 + bf16 pack into the K-major swizzled P tile the PV MMA reads), 64 columns per tcgen05.ld; (m, l)
//                and the 64 output channels of the row live in registers:  O = O * exp2(m_old - m_new) + (P V read back
//                from TMEM one tile late, between the two passes of the next tile, so its MMA latency is never waited for)
// Tile B has one more K/V tile than A (its causal diagonal tile); A's diagonal tile is a full tile for B.
// Window 0's zero look-back keys (reference quirk Q1) enter analytically: m starts at 0 and l at w.
#include "tc_ptx.cuh"
#include "../../include/progen_b200.h"

namespace {

using namespace tc;

constexpr int BQ = 128, BKV = 128, DH = 64;
constexpr int KV_STAGES = 3;
constexpr int Q_BYTES = BQ * DH * 2;              // 16 KiB per query tile
constexpr int K_BYTES = BKV * DH * 2;             // 16 KiB
constexpr int KV_BYTES = 2 * K_BYTES;             // K then V
constexpr int P_BYTES = BQ * BKV * 2;             // 32 KiB per group: two [128 x 64] K-major sub-tiles
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = 2 * Q_BYTES + KV_STAGES * KV_BYTES + 2 * P_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 512;                    // S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384)
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct PairDev {
  int B, n, w, h;
  bf16* out;
  float* lse;
};

// one work item = (batch, head, pair of query tiles).  K/V tile kt of the item: look-back tiles first, then own-window
// tiles 0 .. qa_tile+1.  Group A (g = 0) stops one tile earlier; its last tile and B's last tile are causal diagonals.
struct PairItem { int b, hh, q0, win, nprev, nA; };

__device__ __forceinline__ bool decode_pair(const PairDev& a, int wi, PairItem& it) {
  const int pairs = a.n / (2 * BQ);
  if (wi >= a.B * a.h * pairs) return false;
  const int p = wi % pairs;
  const int r = wi / pairs;
  it.hh = r % a.h;
  it.b = r / a.h;
  it.q0 = p * 2 * BQ;
  it.win = it.q0 / a.w;
  const int i0 = it.q0 % a.w;
  it.nprev = it.win > 0 ? a.w / BKV : 0;
  it.nA = it.nprev + i0 / BKV + 1;                           // tiles seen by A; B sees nA + 1
  return true;
}
__device__ __forceinline__ int pair_key_pos(const PairDev& a, const PairItem& it, int kt) {
  return kt < it.nprev ? (it.win - 1) * a.w + kt * BKV : it.win * a.w + (kt - it.nprev) * BKV;
}

__global__ void __launch_bounds__(384, 1) attn_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const PairDev a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                             // Q_A, Q_B
  const uint32_t sKV = sQ + 2 * Q_BYTES;
  const uint32_t sP = sKV + KV_STAGES * KV_BYTES;            // P_A, P_B
  const uint32_t bars = sP + 2 * P_BYTES;
  const uint32_t q_full = bars, q_empty = bars + 8;
  auto kv_full = [&](int s) { return bars + 16 + 8 * s; };
  auto kv_empty = [&](int s) { return bars + 16 + 8 * (KV_STAGES + s); };
  auto s_full = [&](int g) { return bars + 64 + 8 * g; };
  auto p_full = [&](int g) { return bars + 80 + 8 * g; };
  auto o_full = [&](int g) { return bars + 96 + 8 * g; };
  const uint32_t tmem_slot = bars + 128;
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = a.h * DH;

  if (warp == 0 && lane == 0) prefetch_tensormap(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 4);       // one arrival per softmax warp of the group
      mbar_init(o_full(g), 1);
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
      PairItem it;
      for (int wi = blockIdx.x; decode_pair(a, wi, it); wi += gridDim.x) {
        const int row0 = it.b * a.n;
        mbar_wait(q_empty, q_phase ^ 1);
        mbar_expect_tx(q_full, 2 * Q_BYTES);
        tma_load_2d(sQ, &tmap_qkv, q_full, it.hh * DH, row0 + it.q0);
        tma_load_2d(sQ + Q_BYTES, &tmap_qkv, q_full, it.hh * DH, row0 + it.q0 + BQ);
        q_phase ^= 1;
        for (int kt = 0; kt <= it.nA; ++kt) {
          mbar_wait(kv_empty(stage), kv_phase ^ 1);
          const uint32_t dst = sKV + stage * KV_BYTES;
          const int kp = row0 + pair_key_pos(a, it, kt);
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
      int stage = 0;                                                       // ring position of K/V tile j
      uint32_t kv_phase = 0, q_phase = 0;
      uint32_t pcount[2] = {0, 0};                                         // tiles published per group (p_full parity)
      PairItem it;
      auto issue_qk = [&](int g, int st) {
        tcgen05_fence_after();
        const uint64_t ad = make_smem_desc<false>(sQ + g * Q_BYTES);
        const uint64_t bd = make_smem_desc<false>(sKV + st * KV_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tmem_base + g * BKV, ad + 2 * k, bd + 2 * k, idesc_qk, k > 0);
        tcgen05_commit(s_full(g));
      };
      auto wait_p = [&](int g) {                                           // P_g(j) is in shared memory, S_g is free
        mbar_wait(p_full(g), pcount[g] & 1);
        ++pcount[g];
      };
      auto issue_pv = [&](int g, int st) {
        tcgen05_fence_after();
        const uint32_t pbase = sP + g * P_BYTES;
        const uint64_t vd = make_smem_desc<true>(sKV + st * KV_BYTES + K_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t pd = make_smem_desc<false>(pbase + (k >> 2) * (BQ * 128)) + 2 * (k & 3);
          umma_bf16(tmem_base + 256 + g * DH, pd, vd + (uint64_t)(k * (2048 >> 4)), idesc_pv, k > 0);
        }
        tcgen05_commit(o_full(g));
      };
      for (int wi = blockIdx.x; decode_pair(a, wi, it); wi += gridDim.x) {
        mbar_wait(q_full, q_phase);
        q_phase ^= 1;
        const int nA = it.nA, nB = it.nA + 1;
        mbar_wait(kv_full(stage), kv_phase);
        issue_qk(0, stage);
        issue_qk(1, stage);
        for (int j = 0; j < nB; ++j) {
          int nstage = stage + 1;
          uint32_t nphase = kv_phase;
          if (nstage == KV_STAGES) { nstage = 0; nphase ^= 1; }
          bool next_ready = false;
          // S_g(j+1) goes FIRST: the group needs it to continue, while P_g(j) V_j is only read back after the next row max
          if (j < nA) {
            wait_p(0);
            if (j + 1 < nA) {
              mbar_wait(kv_full(nstage), nphase);
              next_ready = true;
              issue_qk(0, nstage);
            }
            issue_pv(0, stage);
          }
          wait_p(1);
          if (j + 1 < nB) {
            if (!next_ready) mbar_wait(kv_full(nstage), nphase);
            issue_qk(1, nstage);
          } else {
            tcgen05_commit(q_empty);                                       // every QK of this pair has been issued
          }
          issue_pv(1, stage);
          tcgen05_commit(kv_empty(stage));                                 // K_j, V_j free once every MMA so far retires
          stage = nstage;
          kv_phase = nphase;
        }
      }
    }
  } else if (warp >= 4) {
    // ============================================================================ softmax: group g, thread == query row
    const int q = warp & 3;
    const int g = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float sc = 0.125f * LOG2E;                                        // 1/sqrt(64) in log2 units
    const uint32_t s_addr = tmem_base + g * BKV + lane_addr;
    const uint32_t o_addr = tmem_base + 256 + g * DH + lane_addr;
    uint8_t* pbase = gen_base + (sP - smem_base) + g * P_BYTES + row * 128;
    uint32_t tcount = 0;                                                    // tiles processed by this group (barrier parity)
    PairItem it;
    for (int wi = blockIdx.x; decode_pair(a, wi, it); wi += gridDim.x) {
      const int nt = it.nA + g;
      float m_run = it.win == 0 ? 0.f : -INFINITY;                          // quirk Q1: w zero keys with logit 0
      float l_run = it.win == 0 ? (float)a.w : 0.f;
      float o[DH];
#pragma unroll
      for (int i = 0; i < DH; ++i) o[i] = 0.f;
      float corr_pending = 1.f;
      for (int j = 0; j < nt; ++j, ++tcount) {
        const bool diag = j == nt - 1;
        mbar_wait(s_full(g), tcount & 1);
        tcgen05_fence_after();
        // pass 1: row maximum, 64 columns per TMEM load (halves entirely above the diagonal are skipped: warp-uniform)
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // 4 independent chains: a single one is 128 dependent FMNMX
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          if (diag && 2 * hc > q) continue;
          float t[64];
          tmem_ld64(s_addr + hc * 64, t);
          if (diag) {
#pragma unroll
            for (int i = 0; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], hc * 64 + i <= row ? t[i] : -INFINITY);
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], t[i]);
          }
        }
        const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        const float m_new = fmaxf(m_run, mx * sc);
        const float corr = ex2_approx(m_run - m_new);                       // ex2(-inf) = 0 on the first tile
        // P_{j-1} V_{j-1} has had the whole row-max pass to complete: fold it in (also frees P_g for pass 2 below)
        if (j > 0) {
          mbar_wait(o_full(g), (tcount - 1) & 1);
          tcgen05_fence_after();
          float t[64];
          tmem_ld64(o_addr, t);
#pragma unroll
          for (int i = 0; i < DH; ++i) o[i] = o[i] * corr_pending + t[i];
          tcgen05_fence_before();                                           // O_g read before the next PV overwrites it
        }
        corr_pending = corr;
        // pass 2: p = exp2(s c - m), packed to bf16 into the K-major swizzled P tile (sub-tile = 64 keys)
        float rs4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          uint8_t* prow = pbase + hc * (BQ * 128);
          if (diag && 2 * hc > q) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
          float t[64];
          tmem_ld64(s_addr + hc * 64, t);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float p[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float e = ex2_approx(t[ch * 8 + i] * sc - m_new);
              p[i] = (diag && hc * 64 + ch * 8 + i > row) ? 0.f : e;
              rs4[i & 3] += p[i];
            }
            uint4 u;
            u.x = pack_bf16x2(p[0], p[1]); u.y = pack_bf16x2(p[2], p[3]); u.z = pack_bf16x2(p[4], p[5]); u.w = pack_bf16x2(p[6], p[7]);
            *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = u;
          }
        }
        l_run = l_run * corr + ((rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));
        m_run = m_new;
        tcgen05_fence_before();                                             // my reads of S_g precede the next QK into it
        fence_proxy_async();                                                // generic-proxy smem writes -> async proxy (MMA)
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      {                                                                     // the last tile's P V
        mbar_wait(o_full(g), (tcount - 1) & 1);
        tcgen05_fence_after();
        float t[64];
        tmem_ld64(o_addr, t);
#pragma unroll
        for (int i = 0; i < DH; ++i) o[i] = o[i] * corr_pending + t[i];
        tcgen05_fence_before();
      }
      // O / l -> bf16 (this thread's whole 128-byte row of the head); lse in natural-log units
      const long long t = (long long)it.b * a.n + it.q0 + g * BQ + row;
      const float inv = 1.f / l_run;
      bf16* op = a.out + t * I + it.hh * DH;
#pragma unroll
      for (int c = 0; c < DH; c += 8) {
        uint4 u;
        u.x = pack_bf16x2(o[c] * inv, o[c + 1] * inv); u.y = pack_bf16x2(o[c + 2] * inv, o[c + 3] * inv);
        u.z = pack_bf16x2(o[c + 4] * inv, o[c + 5] * inv); u.w = pack_bf16x2(o[c + 6] * inv, o[c + 7] * inv);
        *reinterpret_cast<uint4*>(op + c) = u;
      }
      a.lse[t * a.h + it.hh] = m_run * LN2 + logf(l_run);
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

// Paired forward; returns 1 when the shape is not eligible (caller falls back to the one-tile kernel of attn_tc.cu).
int attn_fwd_pair_launch(const void* qkv, void* out, float* lse, int B, int seq_len, int window, int heads, cudaStream_t stream) {
  static int enabled = [] { const char* e = getenv("PROGEN_ATTN_PAIR"); return e ? atoi(e) : 1; }();
  if (!enabled || window % (2 * BQ) != 0) return 1;
  const long long T = (long long)B * seq_len;
  const int I = heads * DH;
  CUtensorMap tm;
  int rc = pg_tensor_map_2d_bf16(qkv, 3ull * I, (uint64_t)T, 3ull * I, DH, BQ, &tm);
  if (rc) return rc;
  static bool once = false;
  if (!once) {
    PG_CUDA(cudaFuncSetAttribute(attn_fwd_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    once = true;
  }
  PairDev a{B, seq_len, window, heads, (bf16*)out, lse};
  const long long items = (long long)B * heads * (seq_len / (2 * BQ));
  const int grid = (int)(items < pg_num_sms() ? items : pg_num_sms());
  attn_fwd_pair_kernel<<<grid, 384, SMEM_BYTES, stream>>>(tm, a);
  PG_LAUNCH_CHECK();
  return PROGEN_OK;
}
