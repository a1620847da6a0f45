// Sliding-window attention FORWARD, round-2 kernel (window % 256 == 0, dim_head 64): reference progen.py:88-102.
//
// Same work decomposition as attn_tc_pair.cu (a CTA owns the PAIR of adjacent 128-query tiles A, B of one window and
// streams the K|V tiles they share once; thread == query row == TMEM lane), different data path.  ncu on the round-1
// kernel: 16 % tensor pipe, 38 % issue slots, every warp waiting on an mbarrier most of the time — each 128-key step was
// a serial chain of two TMEM passes over S, a read-back of P V, a shared-memory P tile and a proxy fence.  Here:
//
//   * ONE TMEM pass: a thread loads its whole row of S (128 fp32) into registers, takes the maximum, exponentiates and
//     packs in place (setmaxnreg gives the softmax warps 216 registers);
//   * P never touches shared memory: it is written to tensor memory as packed bf16 (tcgen05.st) and the P V product reads
//     its A operand from there (tcgen05.mma, A in TMEM);
//   * O accumulates in TMEM across the steps of an item (accumulate flag), it is rescaled (tcgen05.ld / st) only when a
//     row maximum has grown by more than 2^8 since the maximum the exponents are currently taken against ("lazy
//     rescale"; exact: l and O carry the same offset, the final O / l and the lse do not depend on it);
//   * packed fp32x2 arithmetic (FFMA2 / FADD2), 3-input maximum, and a share of the exponentials on the FMA pipe
//     (cubic 2^f, tc::ex2_poly) — 16 MUFU results / clk / SM are the floor of this kernel, not the tensor pipe;
//   * one MMA-issuing thread per group, Q double-buffered across items, work items ordered heaviest first (the
//     round-robin order of round 1 gave a quarter of the CTAs only window-0 items, i.e. half the work of the others).
//
//   warp 0      : TMA producer (Q_A|Q_B per item, K_j|V_j through a ring)
//   warp 1, 3   : MMA issuer of group A, B:   S_g = Q_g K_j^T (128x128x64)   O_g (+)= P_g V_j (128x64x128, A from TMEM)
//   warp 2      : TMEM allocator
//   warps 4..7  : softmax group A (rows q0 .. q0+127), warps 8..11: group B (q0+128 ..)
// TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512).
// P has its own columns, so S_g(j+1) = Q_g K_{j+1}^T is issued the moment group g has READ S_g(j) into registers and runs
// under the group's own exponentials: the softmax warps never wait for an MMA they have just requested (ncu on the first
// version of this kernel, with P aliased over S: the two groups ran IN PHASE — both exponentiating, then both waiting for
// P V + Q K^T — so the XU pipe and the tensor pipe took turns instead of overlapping: 26 % XU, 19 % tensor).
// Window 0's zero look-back keys (reference quirk Q1) enter analytically: m starts at 0 and l at w.
#include "tc_ptx.cuh"
#include "../../include/progen_b200.h"

namespace {

using namespace tc;

constexpr int BQ = 128, BKV = 128, DH = 64;
constexpr int KV_STAGES = 4;
constexpr int Q_BYTES = BQ * DH * 2;              // 16 KiB per query tile
constexpr int K_BYTES = BKV * DH * 2;             // 16 KiB
constexpr int KV_BYTES = 2 * K_BYTES;             // K then V
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = 4 * Q_BYTES + KV_STAGES * KV_BYTES + BAR_BYTES + 1024;
constexpr int TMEM_COLS = 512;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float RESCALE_THRESHOLD = 8.f;          // log2 units: P stays below 2^8, far inside bf16 / fp32 range

struct FwdDev {
  int B, n, w, h;
  bf16* out;
  float* lse;
};

// one work item = (batch, head, pair of query tiles); K/V tile kt: look-back tiles first, then own-window tiles
// 0 .. qa_tile+1.  Group A (g = 0) stops one tile earlier; its last tile and B's last tile are causal diagonals.
struct Item { int b, hh, q0, win, nprev, nA; };

// heaviest first: every item outside window 0 (look-back tiles present), then window 0's
__device__ __forceinline__ bool decode_item(const FwdDev& a, int wi, Item& it) {
  const int pairs = a.n / (2 * BQ), ppw = a.w / (2 * BQ);
  const int bh = a.B * a.h;
  if (wi >= bh * pairs) return false;
  const int heavy = bh * (pairs - ppw);
  int p, r;
  if (wi < heavy) { p = ppw + wi % (pairs - ppw); r = wi / (pairs - ppw); }
  else { const int v = wi - heavy; p = v % ppw; r = v / ppw; }
  it.hh = r % a.h;
  it.b = r / a.h;
  it.q0 = p * 2 * BQ;
  it.win = it.q0 / a.w;
  const int i0 = it.q0 % a.w;
  it.nprev = it.win > 0 ? a.w / BKV : 0;
  it.nA = it.nprev + i0 / BKV + 1;                           // tiles seen by A; B sees nA + 1
  return true;
}
__device__ __forceinline__ int key_pos(const FwdDev& a, const Item& it, int kt) {
  return kt < it.nprev ? (it.win - 1) * a.w + kt * BKV : it.win * a.w + (kt - it.nprev) * BKV;
}

// POLY: every 4th exponential of a row is evaluated on the FMA pipe instead of the MUFU.
// LOCK: the two softmax warps that share an SM sub-partition (one of each group) take turns in their exponential loops
// (a shared-memory lock per sub-partition), so one group's XU-bound phase runs against the other group's loads / maximum /
// stores instead of both halving each other's MUFU rate at the same time.
template <bool POLY, bool LOCK>
__global__ void __launch_bounds__(384, 1) attn_fwd_ts_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const FwdDev a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                             // [qbuf][g]
  const uint32_t sKV = sQ + 4 * Q_BYTES;
  const uint32_t bars = sKV + KV_STAGES * KV_BYTES;
  auto q_full = [&](int b) { return bars + 8 * b; };
  auto q_empty = [&](int b) { return bars + 16 + 8 * b; };
  auto kv_full = [&](int s) { return bars + 32 + 8 * s; };
  auto kv_empty = [&](int s) { return bars + 64 + 8 * s; };
  auto s_full = [&](int g) { return bars + 96 + 8 * g; };
  auto p_full = [&](int g) { return bars + 112 + 8 * g; };
  auto o_done = [&](int g) { return bars + 128 + 8 * g; };
  auto o_free = [&](int g) { return bars + 144 + 8 * g; };
  auto s_read = [&](int g) { return bars + 160 + 8 * g; };
  const uint32_t tmem_slot = bars + 176;
  volatile int* xu_lock = reinterpret_cast<volatile int*>(smem_raw + (smem_base - smem_u32(smem_raw)) + (bars - smem_base) + 192);   // [4], one per SMSP
  uint8_t* gen_base = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = a.h * DH;

  if (warp == 0 && lane == 0) prefetch_tensormap(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(q_full(b), 1); mbar_init(q_empty(b), 2); }
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full(g), 1);
      mbar_init(p_full(g), 4);       // one arrival per softmax warp of the group
      mbar_init(o_done(g), 1);
      mbar_init(o_free(g), 4);
      mbar_init(s_read(g), 4);
    }
    for (int i = 0; i < 4; ++i) xu_lock[i] = 0;
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - smem_base));

  if (warp < 4) {
    setmaxnreg_dec<64>();
    if (warp == 0) {
      // ============================================================================ TMA producer
      // (the whole warp runs the loop so that addresses / coordinates are warp-uniform; one elected lane issues)
      {
        int stage = 0;
        uint32_t kv_phase = 0, item = 0;
        Item it;
        for (int wi = blockIdx.x; decode_item(a, wi, it); wi += gridDim.x, ++item) {
          const int row0 = it.b * a.n;
          const int qb = item & 1;
          mbar_wait(q_empty(qb), ((item >> 1) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(q_full(qb), 2 * Q_BYTES);
            tma_load_2d(sQ + (2 * qb) * Q_BYTES, &tmap_qkv, q_full(qb), it.hh * DH, row0 + it.q0);
            tma_load_2d(sQ + (2 * qb + 1) * Q_BYTES, &tmap_qkv, q_full(qb), it.hh * DH, row0 + it.q0 + BQ);
          }
          __syncwarp();
          for (int kt = 0; kt <= it.nA; ++kt) {
            mbar_wait(kv_empty(stage), kv_phase ^ 1);
            const uint32_t dst = sKV + stage * KV_BYTES;
            const int kp = row0 + key_pos(a, it, kt);
            if (elect_one()) {
              mbar_expect_tx(kv_full(stage), KV_BYTES);
              tma_load_2d(dst, &tmap_qkv, kv_full(stage), I + it.hh * DH, kp);
              tma_load_2d(dst + K_BYTES, &tmap_qkv, kv_full(stage), 2 * I + it.hh * DH, kp);
            }
            __syncwarp();
            if (++stage == KV_STAGES) { stage = 0; kv_phase ^= 1; }
          }
        }
      }
    } else if (warp == 1 || warp == 3) {
      // ============================================================================ MMA issuer of group g
      // (whole warp: descriptors stay in uniform registers — see tc::elect_one; one elected lane issues tcgen05)
      {
        constexpr uint32_t idesc_qk = make_idesc(BQ, BKV, false, false);     // S[128 x 128] = Q (K-major) x K (K-major)
        constexpr uint32_t idesc_pv = make_idesc(BQ, DH, false, true);       // O[128 x 64] += P (TMEM) x V (MN-major)
        const int g = warp == 1 ? 0 : 1;
        const uint32_t tS = tmem_base + g * BKV, tO = tmem_base + 256 + g * DH, tP = tmem_base + 384 + g * DH;
        int stage = 0;
        uint32_t kv_phase = 0, item = 0, pcount = 0;
        Item it;
        auto issue_qk = [&](int qb, int st) {
          tcgen05_fence_after();
          const uint64_t ad = make_smem_desc<false>(sQ + (2 * qb + g) * Q_BYTES);
          const uint64_t bd = make_smem_desc<false>(sKV + st * KV_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < DH / 16; ++k) umma_bf16(tS, ad + 2 * k, bd + 2 * k, idesc_qk, k > 0);
            tcgen05_commit(s_full(g));
          }
          __syncwarp();
        };
        for (int wi = blockIdx.x; decode_item(a, wi, it); wi += gridDim.x, ++item) {
          const int qb = item & 1;
          const int n_g = it.nA + g, n_all = it.nA + 1;
          mbar_wait(q_full(qb), (item >> 1) & 1);
          mbar_wait(kv_full(stage), kv_phase);
          if (pcount > 0) mbar_wait(s_read(g), (pcount - 1) & 1);              // the previous item's last S_g has been read
          issue_qk(qb, stage);
          for (int j = 0; j < n_all; ++j) {
            int nstage = stage + 1;
            uint32_t nphase = kv_phase;
            if (nstage == KV_STAGES) { nstage = 0; nphase ^= 1; }
            if (j < n_g) {
              if (j + 1 < n_g) {
                mbar_wait(kv_full(nstage), nphase);
                mbar_wait(s_read(g), pcount & 1);                            // S_g(j) is in the group's registers: S_g is free
                issue_qk(qb, nstage);                                        // runs under the group's exponentials of tile j
              } else {
                if (elect_one()) tcgen05_commit(q_empty(qb));                // every Q K^T of this group has been issued
                __syncwarp();
              }
              mbar_wait(p_full(g), pcount & 1);                              // P_g(j) is in tensor memory
              ++pcount;
              if (j == 0 && item > 0) mbar_wait(o_free(g), (item - 1) & 1);    // previous item's O_g has been read out
              tcgen05_fence_after();
              const uint64_t vd = make_smem_desc<true>(sKV + stage * KV_BYTES + K_BYTES);
              if (elect_one()) {
#pragma unroll
                for (int k = 0; k < BKV / 16; ++k)
                  umma_bf16_ts(tO, tP + 8 * k, vd + (uint64_t)(k * (2048 >> 4)), idesc_pv, (j > 0 || k > 0) ? 1u : 0u);
                tcgen05_commit(o_done(g));
                tcgen05_commit(kv_empty(stage));                             // my MMAs on K_j / V_j (the other issuer adds its own)
              }
              __syncwarp();
            } else {
              // a tile only the other group uses: wait until it has landed so the arrival lands in the right phase
              mbar_wait(kv_full(stage), kv_phase);
              if (elect_one()) mbar_arrive(kv_empty(stage));
              __syncwarp();
            }
            stage = nstage;
            kv_phase = nphase;
          }
        }
      }
    }
  } else {
    // ============================================================================ softmax: group g, thread == query row
    setmaxnreg_inc<216>();
    const int q = warp & 3;
    const int g = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const float sc = 0.125f * LOG2E;                                        // 1/sqrt(64) in log2 units
    const uint32_t s_addr = tmem_base + g * BKV + lane_addr;
    const uint32_t o_addr = tmem_base + 256 + g * DH + lane_addr;
    const uint32_t p_addr = tmem_base + 384 + g * DH + lane_addr;
    int* my_lock = const_cast<int*>(xu_lock) + q;
    uint32_t tcount = 0, item = 0;                                          // tiles processed by this group (barrier parity)
    Item it;
    for (int wi = blockIdx.x; decode_item(a, wi, it); wi += gridDim.x, ++item) {
      const int nt = it.nA + g;
      float m_used = it.win == 0 ? 0.f : -INFINITY;                         // quirk Q1: w zero keys with logit 0
      float l_run = it.win == 0 ? (float)a.w : 0.f;
      for (int j = 0; j < nt; ++j, ++tcount) {
        const bool diag = j == nt - 1;
        const int nch = diag ? q + 1 : 4;                                   // 32-column chunks with any visible key (warp-uniform)
        mbar_wait(s_full(g), tcount & 1);
        tcgen05_fence_after();
        uint32_t s[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nch) tmem_ld32_issue(s_addr + c * 32, s[c]);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < nch) tmem_ld32_wait(s[c]);                                // (names the registers: no use may move above it)
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_read(g));                              // S_g is in registers: the next Q K^T may overwrite it
        // row maximum (chunks above the diagonal are skipped, the diagonal chunk is masked per element)
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
            if (diag && c == q) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i > lane) s[c][i] = 0xff800000u;                       // -inf: key after query
            }
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              mx0 = fmax3(mx0, __uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1]));
              mx1 = fmax3(mx1, __uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3]));
              mx2 = fmax3(mx2, __uint_as_float(s[c][i + 4]), __uint_as_float(s[c][i + 5]));
              mx3 = fmax3(mx3, __uint_as_float(s[c][i + 6]), __uint_as_float(s[c][i + 7]));
            }
          }
        }
        const float m_cand = fmaxf(m_used, fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc);
        const bool need = m_cand - m_used > RESCALE_THRESHOLD;              // also true on the first tile (m_used = -inf)
        bool pv_waited = false;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? m_cand : m_used;
          const float corr = ex2f(m_used - m_new);                          // ex2(-inf) = 0 on the first tile
          l_run *= corr;
          m_used = m_new;
          if (j > 0) {
            mbar_wait(o_done(g), (tcount - 1) & 1);                         // P V of the previous step has retired
            tcgen05_fence_after();
            pv_waited = true;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {                                // 32 columns at a time: S stays live in registers
              uint32_t o[32];
              tmem_ld32_issue(o_addr + h2 * 32, o);
              tmem_ld32_wait(o);
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
              tmem_st<32>(o_addr + h2 * 32, o);
            }
          }
        }
        // p = exp2(s c - m), packed to bf16 pairs; written to P_g once the previous P V (which reads P_g) has retired
        const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_used, -m_used);
        float2 rs0 = make_float2(0.f, 0.f), rs1 = make_float2(0.f, 0.f);
        uint32_t pk[4][16];
        if (LOCK) {
          if (lane == 0) while (atomicExch(my_lock, 1) != 0) __nanosleep(32);
          __syncwarp();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < nch) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              float2 x0 = ffma2(make_float2(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])), sc2, nm2);
              float2 x1 = ffma2(make_float2(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])), sc2, nm2);
              x0.x = ex2f(x0.x);
              x0.y = ex2f(x0.y);
              x1.x = ex2f(x1.x);
              x1.y = POLY ? ex2_poly(x1.y) : ex2f(x1.y);
              rs0 = fadd2(rs0, x0);
              rs1 = fadd2(rs1, x1);
              pk[c][i / 2] = pack_bf16x2(x0.x, x0.y);
              pk[c][i / 2 + 1] = pack_bf16x2(x1.x, x1.y);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[c][i] = 0u;
          }
        }
        if (LOCK) {
          __syncwarp();
          if (lane == 0) atomicExch(my_lock, 0);
        }
        if (j > 0 && !pv_waited) {
          mbar_wait(o_done(g), (tcount - 1) & 1);                           // P V of the previous step no longer reads P_g
          tcgen05_fence_after();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st<16>(p_addr + c * 16, pk[c]);
        l_run += (rs0.x + rs0.y) + (rs1.x + rs1.y);
        tmem_st_wait();
        tcgen05_fence_before();                                             // my TMEM reads / writes precede the MMAs that follow
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(g));
      }
      // the item's O: wait for the last P V, read it out, free O_g for the next item
      mbar_wait(o_done(g), (tcount - 1) & 1);
      tcgen05_fence_after();
      uint32_t o[2][32];
      tmem_ld32_issue(o_addr, o[0]);
      tmem_ld32_issue(o_addr + 32, o[1]);
      tmem_ld32_wait(o[0]);
      tmem_ld32_wait(o[1]);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free(g));
      // O / l -> bf16 (this thread's whole 128-byte row of the head); lse in natural-log units
      const long long t = (long long)it.b * a.n + it.q0 + g * BQ + row;
      const float inv = 1.f / l_run;
      bf16* op = a.out + t * I + it.hh * DH;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
        for (int c = 0; c < 32; c += 8) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[h2][c]) * inv, __uint_as_float(o[h2][c + 1]) * inv);
          u.y = pack_bf16x2(__uint_as_float(o[h2][c + 2]) * inv, __uint_as_float(o[h2][c + 3]) * inv);
          u.z = pack_bf16x2(__uint_as_float(o[h2][c + 4]) * inv, __uint_as_float(o[h2][c + 5]) * inv);
          u.w = pack_bf16x2(__uint_as_float(o[h2][c + 6]) * inv, __uint_as_float(o[h2][c + 7]) * inv);
          *reinterpret_cast<uint4*>(op + h2 * 32 + c) = u;
        }
      }
      a.lse[t * a.h + it.hh] = m_used * LN2 + logf(l_run);
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

// Round-2 forward; returns 1 when the shape is not eligible (caller falls back to the round-1 kernels).
int attn_fwd_ts_launch(const void* qkv, void* out, float* lse, int B, int seq_len, int window, int heads, cudaStream_t stream) {
  static int mode = [] { const char* e = getenv("PROGEN_ATTN_TS"); return e ? atoi(e) : 2; }();   // 0 off, 1 MUFU only, 2 + FMA-pipe exp2
  if (!mode || window % (2 * BQ) != 0) return 1;
  const long long T = (long long)B * seq_len;
  const int I = heads * DH;
  CUtensorMap tm;
  int rc = pg_tensor_map_2d_bf16(qkv, 3ull * I, (uint64_t)T, 3ull * I, DH, BQ, &tm);
  if (rc) return rc;
  static bool once = false;
  if (!once) {
    PG_CUDA(cudaFuncSetAttribute(attn_fwd_ts_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    PG_CUDA(cudaFuncSetAttribute(attn_fwd_ts_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    PG_CUDA(cudaFuncSetAttribute(attn_fwd_ts_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    PG_CUDA(cudaFuncSetAttribute(attn_fwd_ts_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    once = true;
  }
  FwdDev a{B, seq_len, window, heads, (bf16*)out, lse};
  const long long items = (long long)B * heads * (seq_len / (2 * BQ));
  const int grid = (int)(items < pg_num_sms() ? items : pg_num_sms());
  // mode: 1 MUFU only, 2 + FMA-pipe exp2 share, 3 = 1 + XU lock, 4 = 2 + XU lock
  if (mode == 4) attn_fwd_ts_kernel<true, true><<<grid, 384, SMEM_BYTES, stream>>>(tm, a);
  else if (mode == 3) attn_fwd_ts_kernel<false, true><<<grid, 384, SMEM_BYTES, stream>>>(tm, a);
  else if (mode == 2) attn_fwd_ts_kernel<true, false><<<grid, 384, SMEM_BYTES, stream>>>(tm, a);
  else attn_fwd_ts_kernel<false, false><<<grid, 384, SMEM_BYTES, stream>>>(tm, a);
  PG_LAUNCH_CHECK();
  return PROGEN_OK;
}
