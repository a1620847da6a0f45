// CTA-pair tcgen05 GEMM (cta_group::2): D[M,N] = A[M,K] * B[N,K]^T for the un-batched GEMMs of the model: the activation
// GEMMs (A K-major, fused epilogues) and the weight gradients dW += X^T dY (both operands MN-major, K = tokens, split
// over K with a TMA reduce-add epilogue).
//
// Mainloop — why a CTA pair: with 128x256 tiles the K = 512 GEMMs of this model are bound by L2 -> SM operand traffic
// (every tile pulls 16 KiB of A + 32 KiB of B per k-block; measured ~10-13 TB/s aggregate).  Two CTAs of a cluster (one
// TPC) share a 256 x 256 output tile: each loads its 128 rows of A and only HALF of the B tile (16 + 16 KiB per k-block,
// -33 % L2 traffic per FLOP); one tcgen05.mma.cta_group::2 (M = 256) issued by the leader CTA reads both halves and
// writes each CTA's 128 accumulator rows into that CTA's TMEM.
//
// Epilogue — all global traffic goes through TMA boxes, none through per-thread LDG/STG:
//   The fused epilogues move 0.3-1.1 GB per launch (saved pre-activations in, activations / gradients out).  With
//   per-thread loads the bytes in flight are capped by registers (one 4 KiB chunk per warp): the GLU-backward GEMM ran at
//   313 us against a 176 us HBM floor, and the transposition through shared memory kept the LSU pipe 59-75 % busy.  Here
//   each group of 4 epilogue warps (128 threads = the CTA's 128 accumulator rows; two groups split the 256 columns)
//   owns a ring of shared-memory SLOTS, one [128 rows x <=128 B] box (+ a second box for the two-output kinds):
//     - the group's elected thread TMA-loads the second operand (residual / pre-activations) of chunk i+NB-1 into a
//       slot while chunk i is processed: up to 2 x 16 KiB in flight per group, independent of registers;
//     - a thread reads ITS row from the slot (TMEM lane == row; the TMA swizzle makes lane-per-row 16-byte accesses
//       bank-conflict free), combines it with 32 accumulator columns from TMEM and writes the result row back in place;
//     - fence.proxy.async + 128-thread named barrier, then the elected thread TMA-stores the box (rows beyond M are
//       clipped by the tensor map) and recycles the slot once the store has read it (cp.async.bulk.wait_group.read).
//
// Protocol (r = cluster rank, leader = rank 0):
//   full[s]   (leader's, count 1): leader producer arrives with expect_tx = bytes of BOTH CTAs; both CTAs' TMA loads
//             complete_tx on the leader's barrier (cp.async.bulk.tensor ... .cta_group::2 with the leader's address)
//   empty[s]  (one per CTA, count 1): tcgen05.commit.cta_group::2 ... multicast::cluster -> both producers
//   tfull[a]  (one per CTA, count 1): commit multicast -> both CTAs' epilogue warps
//   tempty[a] (leader's, count 16): 8 local + 8 remote (mapa) epilogue-warp arrivals
//   sready[g][b] (per CTA, count 1): slot b of epilogue group g holds chunk i's second operand / is free for chunk i
#include "tc_ptx.cuh"
#include "gemm.h"

namespace {

using namespace tc;
using namespace tc2;

constexpr int BM = 128;                 // rows per CTA (256 per pair)
constexpr int BN = 256;                 // columns per pair; each CTA stages BN/2 rows of B
constexpr int BK = 64;
constexpr int EW = 8;                   // epilogue warps: 2 groups x 4 warps
constexpr int NGROUP = 2;
constexpr int CPG = BN / NGROUP / 32;   // 32-column chunks per group per tile
constexpr int THREADS = 128 + 32 * EW;
constexpr int A_BYTES = BM * BK * 2;            // 16 KiB
constexpr int B_BYTES = (BN / 2) * BK * 2;      // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int BAR_BYTES = 256;
constexpr int TMEM_COLS = 2 * BN;
constexpr int SMEM_LIMIT = 227 * 1024;

// Per-kind shape of the epilogue slots.  ROWB0: bytes per row of the primary box (the in-place operand for the kinds
// with a second input, else the first output); ROWB1: bytes per row of the second output box (GLU / GELU forward).
template <int KIND, typename TO> struct Epi2 {
  static constexpr bool AUX = epi_has_aux<KIND>;
  static constexpr int ROWB0 = (KIND == EPI_RESIDUAL || KIND == EPI_ACCUM) ? 32 * 4 : KIND == EPI_GLU_BWD ? 64 * (int)sizeof(TO) : 32 * (int)sizeof(TO);
  static constexpr int ROWB1 = KIND == EPI_GLU ? 16 * (int)sizeof(TO) : KIND == EPI_GELU ? 32 * (int)sizeof(TO) : 0;
  static constexpr int NB = AUX ? 3 : 2;                       // slots per group
  static constexpr int SLOT_BYTES = BM * (ROWB0 + ROWB1);      // multiples of 4 KiB: every box stays 1024-byte aligned
  static constexpr int SLOTS_TOTAL = NGROUP * NB * SLOT_BYTES;
  static constexpr int ROT_BYTES = KIND == EPI_ROTARY ? EW * STAGE_WARP_BYTES : 0;    // staged sin/cos loads
  static constexpr int FIXED = 1024 + SLOTS_TOTAL + BAR_BYTES + ROT_BYTES;
  static constexpr int STAGES_FIT = (SMEM_LIMIT - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 6 ? 6 : STAGES_FIT;
  static constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + FIXED;
  static_assert(ROWB0 <= 128 && ROWB0 >= 32 && STAGES >= 3, "epilogue slot shape");
};

struct Gemm2Dev {
  int M, N, K;
  int split_k;                          // work item = (output tile, K slice); > 1 only with the reduce-add epilogue
  EpiArgs epi;
};

// work item t -> output tile origin and k-block range.  Consecutive items are DIFFERENT tiles (the K slices of one
// tile are `tiles` items apart), so concurrently finishing CTAs reduce into different addresses.
// A work item covers 256 rows x (CL * 256) columns: with CL == 2 the two pairs of a 4-CTA cluster take adjacent column
// tiles of the SAME rows (the caller adds pair_index * BN to n0), so the A tile is loaded once and multicast.
template <int CL>
__device__ __forceinline__ bool decode_tile2(const Gemm2Dev& g, int t, int& m0, int& n0, int& kb0, int& kb1) {
  const int m_tiles = (g.M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = g.N / (BN * CL);
  const int tiles = m_tiles * n_tiles;
  if (t >= tiles * g.split_k) return false;
  const int tile = t % tiles, sp = t / tiles;
  m0 = (tile / n_tiles) * (2 * BM);
  n0 = (tile % n_tiles) * (BN * CL);
  const long long kb_total = g.K / BK;
  kb0 = (int)(kb_total * sp / g.split_k);
  kb1 = (int)(kb_total * (sp + 1) / g.split_k);
  return true;
}

// 16-byte piece q of row r inside a [128 x ROWB] box stored with the TMA swizzle of span ROWB (32 / 64 / 128 bytes)
template <int ROWB> __device__ __forceinline__ uint32_t box_off(int r, int q) { return (uint32_t)stage_off<ROWB / 16>(r, q); }

// read / write N values of row r from / to a swizzled box (TMEM lane == row: one row per thread)
template <int ROWB, typename T, int N> __device__ __forceinline__ void box_read(const uint8_t* box, int r, float (&v)[N]) {
  constexpr int EPP = 16 / (int)sizeof(T);
  static_assert(N * (int)sizeof(T) == ROWB, "row width");
#pragma unroll
  for (int q = 0; q < ROWB / 16; ++q)
    WarpStagedIO::unpack16<T>(*reinterpret_cast<const uint4*>(box + box_off<ROWB>(r, q)), &v[q * EPP]);
}
template <int ROWB, typename T, int N> __device__ __forceinline__ void box_write(uint8_t* box, int r, const float (&v)[N]) {
  constexpr int EPP = 16 / (int)sizeof(T);
  static_assert(N * (int)sizeof(T) == ROWB, "row width");
#pragma unroll
  for (int q = 0; q < ROWB / 16; ++q)
    *reinterpret_cast<uint4*>(box + box_off<ROWB>(r, q)) = WarpStagedIO::pack16<T>(&v[q * EPP]);
}

template <bool A_MN, bool B_MN, int KIND, typename TO, int CL>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                const __grid_constant__ CUtensorMap tma_aux, const __grid_constant__ CUtensorMap tma_out,
                const __grid_constant__ CUtensorMap tma_out2, const Gemm2Dev g) {
  using E = Epi2<KIND, TO>;
  constexpr int STAGES = E::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t slots_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = slots_base + E::SLOTS_TOTAL;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (6 + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (12 + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (14 + s); };
  auto sready_bar = [&](int grp, int b) { return bar_base + 8u * (16 + grp * 3 + b); };
  const uint32_t tmem_slot = bar_base + 8u * 22;
  uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // cluster = CL pairs; crank = rank in the cluster, rank = rank inside the pair, pidx = which pair (column tile)
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1u, pidx = crank >> 1, lead_crank = crank & ~1u;
  const bool leader = rank == 0;
  const int pair = blockIdx.x / (2 * CL), npairs = gridDim.x / (2 * CL);      // work-item stream: one per CLUSTER
  const int ncol = (int)pidx * BN;                                            // my pair's column offset inside the item
  constexpr uint16_t ALL_CTAS = (uint16_t)((1u << (2 * CL)) - 1u);

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tma_a); prefetch_tensormap(&tma_b); prefetch_tensormap(&tma_out);
    if constexpr (E::AUX) prefetch_tensormap(&tma_aux);
    if constexpr (E::ROWB1 > 0) prefetch_tensormap(&tma_out2);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL); }   // empty: every pair's commit
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * EW); }
    for (int grp = 0; grp < NGROUP; ++grp)
      for (int b = 0; b < E::NB; ++b) mbar_init(sready_bar(grp, b), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  cluster_sync();                                   // barriers of BOTH CTAs are initialised before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    {                                                   // whole warp (uniform operands), one elected lane issues
      int stage = 0;
      uint32_t phase = 0;
      int m0, n0, kb0, kb1;
      for (int t = pair; decode_tile2<CL>(g, t, m0, n0, kb0, kb1); t += npairs) {
        n0 += ncol;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint32_t lead_full = mapa(full_bar(stage), lead_crank);
          const int k0 = kb * BK;
          if (elect_one()) {
          if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
          if constexpr (CL == 1) {
            if constexpr (!A_MN) {
              tma_load_2d_pair(sa, &tma_a, lead_full, k0, m0 + (int)rank * BM);                 // my 128 rows of A
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)                                                 // my 2 x 64 columns of A^T
                tma_load_2d_pair(sa + i * 8192, &tma_a, lead_full, m0 + (int)rank * BM + 64 * i, k0);
            }
          } else if (pidx == 0) {
            // 4-CTA cluster: pair 0 loads the A rows both pairs use and multicasts them to the same-rank CTA of pair 1
            const uint16_t mask = (uint16_t)((1u << rank) | (1u << (rank + 2)));
            if constexpr (!A_MN) {
              tma_load_2d_pair_mc(sa, &tma_a, full_bar(stage), mask, k0, m0 + (int)rank * BM);
            } else {
#pragma unroll
              for (int i = 0; i < BM / 64; ++i)
                tma_load_2d_pair_mc(sa + i * 8192, &tma_a, full_bar(stage), mask, m0 + (int)rank * BM + 64 * i, k0);
            }
          }
          if constexpr (!B_MN) {
            tma_load_2d_pair(sb, &tma_b, lead_full, k0, n0 + (int)rank * (BN / 2));             // my 128 rows of B
          } else {
#pragma unroll
            for (int i = 0; i < BN / 2 / 64; ++i)                                               // my 2 x 64 columns of B
              tma_load_2d_pair(sb + i * 8192, &tma_b, lead_full, n0 + (int)rank * (BN / 2) + 64 * i, k0);
          }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    // the WHOLE warp runs the loop (all values warp-uniform -> uniform registers, no per-instruction waterfall); one
    // elected lane issues the tcgen05 instructions
    if (leader) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN, A_MN, B_MN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      int m0, n0, kb0, kb1;
      for (int t = pair; decode_tile2<CL>(g, t, m0, n0, kb0, kb1); t += npairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);              // both CTAs' epilogues drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);                    // both CTAs' operand tiles have landed
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = make_smem_desc<A_MN>(sa);
          const uint64_t bdesc = make_smem_desc<B_MN>(sa + A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * (A_MN ? (2048 >> 4) : 2)),
                             bdesc + (uint64_t)(k * (B_MN ? (2048 >> 4) : 2)), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            tcgen05_commit_pair(empty_bar(stage), ALL_CTAS);    // frees the stage in every CTA of the cluster
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tcgen05_commit_pair(tfull_bar(acc), (uint16_t)(3u << (2 * pidx)));   // accumulators of both CTAs of MY pair are complete
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (each CTA drains its own 128 rows)
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int r_in_tile = q * 32 + lane;
    const bool elected = q == 0 && lane == 0;
    const int m_tiles = (g.M + 2 * BM - 1) / (2 * BM), n_tiles = g.N / (BN * CL);
    const int total_tiles = m_tiles * n_tiles * g.split_k;
    const int my_tiles = total_tiles > pair ? (total_tiles - pair + npairs - 1) / npairs : 0;
    const int nchunks = my_tiles * CPG;
    // chunk i of this group -> (row of the CTA's 128-row block, first accumulator column)
    auto chunk_coords = [&](int i, int& row0, int& col) {
      int m0, n0, kb0, kb1;
      decode_tile2<CL>(g, pair + (i / CPG) * npairs, m0, n0, kb0, kb1);
      row0 = m0 + (int)rank * BM;
      col = n0 + ncol + (grp * CPG + (i % CPG)) * 32;
    };
    auto slot_addr = [&](int b) { return slots_base + (uint32_t)((grp * E::NB + b) * E::SLOT_BYTES); };
    // elected thread: make slot (i % NB) ready for chunk i — TMA-load the second operand, or just mark the slot free
    auto prepare = [&](int i) {
      const int b = i % E::NB;
      if constexpr (E::AUX) {
        int row0, col;
        chunk_coords(i, row0, col);
        mbar_expect_tx(sready_bar(grp, b), BM * E::ROWB0);
        tma_load_2d(slot_addr(b), &tma_aux, sready_bar(grp, b), KIND == EPI_GLU_BWD ? 2 * col : col, row0);
      } else {
        mbar_arrive(sready_bar(grp, b));
      }
    };
    if (elected)
      for (int i = 0; i < E::NB - 1 && i < nchunks; ++i) prepare(i);

    WarpStagedIO rio;                                   // ROTARY only: staged loads of the sin / cos rows
    rio.buf = gen + (bar_base - smem_base) + BAR_BYTES + (warp - 4) * STAGE_WARP_BYTES;
    rio.lane = lane;
    rio.valid_mask = 0xffffffffu;
    float rs[KIND == EPI_ROTARY ? 32 : 1], rc[KIND == EPI_ROTARY ? 32 : 1];

    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t taddr = 0;
    uint32_t vr[32];                                    // accumulator chunk in flight (raw TMEM words)
    // prefetch of the next chunk: A/B on one box -0.8 % (GLU, GLU backward), 0 (residual), +3 % (rotary: register pressure)
    constexpr bool PF = KIND != EPI_ROTARY;
    long long row = 0;
#pragma unroll 1
    for (int i = 0; i < nchunks; ++i) {
      int row0, col;
      chunk_coords(i, row0, col);
      const int ci = i % CPG;
      if (ci == 0) {
        row = row0 + r_in_tile;
        if constexpr (KIND == EPI_ROTARY) {             // this row's 32 (sin, cos) pairs: dim_head == 64 (launcher)
          const long long pos = row % g.epi.seq_len;
          rio.template load<32>(g.epi.rot_sin + pos * 32, 32, rs, true);
          rio.template load<32>(g.epi.rot_cos + pos * 32, 32, rc, true);
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tcgen05_fence_after();
        taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16) + grp * (CPG * 32);
        tmem_ld32_issue(taddr, vr);
      }
      else if (!PF) tmem_ld32_issue(taddr + ci * 32, vr);
      const int b = i % E::NB;
      tmem_ld32_wait(vr);                               // chunk ci (issued at the end of the previous chunk, or just above)
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(vr[j]);
      if (ci == CPG - 1) {                              // accumulator stage fully read by this warp
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(acc), lead_crank));   // my pair leader's barrier (local or remote)
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      mbar_wait(sready_bar(grp, b), (uint32_t)(i / E::NB) & 1u);
      uint8_t* box0 = gen + (slot_addr(b) - smem_base);
      uint8_t* box1 = box0 + BM * E::ROWB0;
      constexpr bool FAST = sizeof(TO) == 2;
      if constexpr (KIND == EPI_STORE) {
        if (g.epi.bias) add_bias<32>(g.epi.bias, col, v);
        box_write<E::ROWB0, TO, 32>(box0, r_in_tile, v);
      } else if constexpr (KIND == EPI_ROTARY) {
        float o[32];
        if ((col & 32) == 0) {                          // which half of the 64-wide head this chunk covers (static indices)
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            o[j] = v[j] * rc[j >> 1] - v[j + 1] * rs[j >> 1];
            o[j + 1] = v[j + 1] * rc[j >> 1] + v[j] * rs[j >> 1];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            o[j] = v[j] * rc[16 + (j >> 1)] - v[j + 1] * rs[16 + (j >> 1)];
            o[j + 1] = v[j + 1] * rc[16 + (j >> 1)] + v[j] * rs[16 + (j >> 1)];
          }
        }
        box_write<E::ROWB0, TO, 32>(box0, r_in_tile, o);
      } else if constexpr (KIND == EPI_RESIDUAL) {
        float r[32];
        box_read<E::ROWB0, float, 32>(box0, r_in_tile, r);
        if (g.epi.bias) add_bias<32>(g.epi.bias, col, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] += v[j];
        box_write<E::ROWB0, float, 32>(box0, r_in_tile, r);
      } else if constexpr (KIND == EPI_GLU) {
        add_bias<32>(g.epi.bias, col, v);
        box_write<E::ROWB0, TO, 32>(box0, r_in_tile, v);
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = v[2 * j] * gelu_fwd<FAST>(v[2 * j + 1]);
        box_write<E::ROWB1, TO, 16>(box1, r_in_tile, o);
      } else if constexpr (KIND == EPI_GELU) {
        add_bias<32>(g.epi.bias, col, v);
        box_write<E::ROWB0, TO, 32>(box0, r_in_tile, v);
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = gelu_fwd<FAST>(v[j]);
        box_write<E::ROWB1, TO, 32>(box1, r_in_tile, o);
      } else if constexpr (KIND == EPI_GLU_BWD) {
        float u[64];
        box_read<E::ROWB0, TO, 64>(box0, r_in_tile, u);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float dh = v[j], val = u[2 * j], gate = u[2 * j + 1];
          float gf, gd;
          gelu_fwd_bwd<FAST>(gate, gf, gd);
          u[2 * j] = dh * gf;
          u[2 * j + 1] = dh * val * gd;
        }
        box_write<E::ROWB0, TO, 64>(box0, r_in_tile, u);
      } else if constexpr (KIND == EPI_ACCUM) {
        box_write<E::ROWB0, float, 32>(box0, r_in_tile, v);
      } else if constexpr (KIND == EPI_GELU_BWD) {
        float u[32];
        box_read<E::ROWB0, TO, 32>(box0, r_in_tile, u);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= gelu_bwd<FAST>(u[j]);
        box_write<E::ROWB0, TO, 32>(box0, r_in_tile, v);
      }
      // the next chunk's accumulators travel TMEM -> registers while this chunk is fenced, synchronised and stored
      if (PF && ci + 1 < CPG) tmem_ld32_issue(taddr + (ci + 1) * 32, vr);
      fence_proxy_async();                              // my shared-memory writes -> visible to the TMA store
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (elected) {
        if constexpr (KIND == EPI_GLU) {
          tma_store_2d(&tma_out2, slot_addr(b), col, row0);
          tma_store_2d(&tma_out, slot_addr(b) + BM * E::ROWB0, col >> 1, row0);
        } else if constexpr (KIND == EPI_GELU) {
          tma_store_2d(&tma_out2, slot_addr(b), col, row0);
          tma_store_2d(&tma_out, slot_addr(b) + BM * E::ROWB0, col, row0);
        } else if constexpr (KIND == EPI_ACCUM) {
          tma_reduce_add_2d(&tma_out, slot_addr(b), col, row0);      // out += acc (fp32 add performed by the L2)
        } else {
          tma_store_2d(&tma_out, slot_addr(b), KIND == EPI_GLU_BWD ? 2 * col : col, row0);
        }
        bulk_commit_group();
        if (i + E::NB - 1 < nchunks) {
          bulk_wait_group_read<1>();                    // the store of chunk i-1 has read its slot: reuse it
          prepare(i + E::NB - 1);
        }
      }
    }
    if (elected) bulk_wait_group<0>();                  // every store has completed before shared memory goes away
  }

  tcgen05_fence_before();
  cluster_sync();                                   // nobody exits (or frees TMEM) while the partner may still touch it
  if (warp == 2) { tcgen05_fence_after(); tmem_dealloc_pair<TMEM_COLS>(tmem_base); }
}

// how many clusters of 2*CL CTAs (1 CTA per SM, full shared memory) the device keeps resident at once: the persistent
// grid must not exceed it, or a late cluster would serialise behind the statically partitioned others
template <bool A_MN, bool B_MN, int KIND, typename TO, int CL>
int max_clusters() {
  using E = Epi2<KIND, TO>;
  static int cached = -1;
  if (cached >= 0) return cached;
  auto kern = gemm_tc2_kernel<A_MN, B_MN, KIND, TO, CL>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, E::SMEM_TOTAL) != cudaSuccess) return cached = 0;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * CL * (pg_num_sms() / (2 * CL)));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = E::SMEM_TOTAL;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  const int cap = pg_num_sms() / (2 * CL);
  if (getenv("PROGEN_DEBUG")) fprintf(stderr, "[progen] gemm_tc2 kind %d: %d resident clusters of %d CTAs (cap %d)\n", KIND, n, 2 * CL, cap);
  return cached = (n < cap ? n : cap);
}

template <bool A_MN, bool B_MN, int KIND, typename TO, int CL>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& taux, const CUtensorMap& tout,
            const CUtensorMap& tout2, const Gemm2Dev& gd, int items, cudaStream_t stream) {
  using E = Epi2<KIND, TO>;
  auto kern = gemm_tc2_kernel<A_MN, B_MN, KIND, TO, CL>;
  int clusters = max_clusters<A_MN, B_MN, KIND, TO, CL>();
  if (clusters <= 0) { progen_set_error("gemm_tc2: no resident cluster of %d CTAs", 2 * CL); return PROGEN_ERR_CUDA; }
  if (items < clusters) clusters = items;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * CL * clusters);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = E::SMEM_TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PG_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, taux, tout, tout2, gd));
  PG_LAUNCH_CHECK();
  return PROGEN_OK;
}

// CL = 2 (4-CTA cluster, A multicast to two pairs) for the mainloop-bound kinds when N holds whole 512-column items
template <bool A_MN, bool B_MN, int KIND, typename TO>
int launch2_auto(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& taux, const CUtensorMap& tout,
                 const CUtensorMap& tout2, Gemm2Dev gd, const GemmArgs& a, cudaStream_t stream) {
  // Measured (profiles/r01_gemm_bench_4cta.txt): correct, but 3-6 % SLOWER than pairs on the config-2 shapes — clusters of
  // 4 need two free TPCs in one GPC, so fewer SMs are usable than with pairs, which costs more than the multicast saves.
  static int quad = [] { const char* e = getenv("PROGEN_GEMM_4CTA"); return e ? atoi(e) : 0; }();
  const int m_tiles = (a.M + 2 * BM - 1) / (2 * BM);
  const bool use4 = quad && a.N % (2 * BN) == 0 && max_clusters<A_MN, B_MN, KIND, TO, 2>() > 0;
  const int cl = use4 ? 2 : 1;
  const int tiles = m_tiles * (a.N / (BN * cl));
  const int streams = use4 ? max_clusters<A_MN, B_MN, KIND, TO, 2>() : max_clusters<A_MN, B_MN, KIND, TO, 1>();
  int split = 1;
  if (a.epi_kind == EPI_ACCUM && a.epi.atomic && streams > 0) {
    // K slices per tile: fill the resident clusters as evenly as possible, keep >= 16 k-blocks per slice, prefer fewer slices
    const int kb_total = a.K / BK;
    double best = 0.0;
    for (int s = 1; s <= 48 && kb_total / s >= 16; ++s) {
      const int items = tiles * s, waves = (items + streams - 1) / streams;
      const double eff = (double)items / (waves * streams);
      if (eff > best + 0.02) { best = eff; split = s; }
    }
  }
  gd.split_k = split;
  if (use4) return launch2<A_MN, B_MN, KIND, TO, 2>(ta, tb, taux, tout, tout2, gd, tiles * split, stream);
  return launch2<A_MN, B_MN, KIND, TO, 1>(ta, tb, taux, tout, tout2, gd, tiles * split, stream);
}

}  // namespace

bool gemm_tc2_eligible(const GemmArgs& a) {
  static int enabled = [] { const char* e = getenv("PROGEN_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  if (!enabled) return false;
  if (a.in_dtype != PG_BF16 || a.batch != 1 || a.causal || a.batch_reduce) return false;
  if (a.N % BN != 0 || a.K % BK != 0 || a.M < 2 * BM) return false;
  const bool am = a.a_mn_major != 0, bm = a.b_mn_major != 0, obf = a.out_dtype == PG_BF16;
  if (a.epi_kind == EPI_ACCUM) {
    // weight gradients dW[M,N] += X^T dY: both operands MN-major, K = tokens.  The K split is re-chosen for 74 CTA pairs
    // (the caller's split_k targets 148 single CTAs); several slices of one tile need the reduce-add to be allowed.
    if (!am || !bm || a.epi.tril || a.M % 8 != 0) return false;
    if (a.split_k > 1 && !a.epi.atomic) return false;
    return a.epi.out && ((uintptr_t)a.epi.out & 15) == 0 && (a.epi.ldo * 4) % 16 == 0;
  }
  if (am || a.split_k != 1) return false;
  const int osz = obf ? 2 : 4;
  // every epilogue pointer / leading dimension must be TMA-addressable (16-byte aligned base and row pitch)
  auto tma_ok = [](const void* p, long long ld, int esz) { return p && ((uintptr_t)p & 15) == 0 && (ld * esz) % 16 == 0; };
  switch (a.epi_kind) {                               // exactly the combinations instantiated below
    case EPI_STORE: return tma_ok(a.epi.out, a.epi.ldo, osz);
    case EPI_ROTARY: return bm && obf && a.epi.dim_head == 64 && a.epi.seq_len % 128 == 0 && tma_ok(a.epi.out, a.epi.ldo, 2);
    case EPI_GLU: case EPI_GELU:
      return bm && obf && a.epi.bias && tma_ok(a.epi.out, a.epi.ldo, 2) && tma_ok(a.epi.out2, a.epi.ldo2, 2);
    case EPI_RESIDUAL:
      return bm && tma_ok(a.epi.out, a.epi.ldo, 4) && (!a.epi.aux || tma_ok(a.epi.aux, a.epi.ldaux, 4));
    case EPI_GLU_BWD: case EPI_GELU_BWD:
      return !bm && obf && tma_ok(a.epi.out, a.epi.ldo, 2) && tma_ok(a.epi.aux, a.epi.ldaux, 2);
    default: return false;
  }
}

int gemm_tc2_launch(const GemmArgs& a, cudaStream_t stream) {
  PG_CHECK_ARG(gemm_tc2_eligible(a));
  PG_CHECK_ARG(a.lda % 8 == 0 && a.ldb % 8 == 0);
  CUtensorMap ta, tb, taux, tout, tout2;
  int rc;
  if (!a.a_mn_major) rc = pg_tensor_map_2d_bf16(a.A, a.K, a.M, a.lda, BK, BM, &ta);
  else               rc = pg_tensor_map_2d_bf16(a.A, a.M, a.K, a.lda, 64, BK, &ta);
  if (rc) return rc;
  if (!a.b_mn_major) rc = pg_tensor_map_2d_bf16(a.B, a.K, a.N, a.ldb, BK, BN / 2, &tb);
  else               rc = pg_tensor_map_2d_bf16(a.B, a.N, a.K, a.ldb, 64, BK, &tb);
  if (rc) return rc;
  const bool bm = a.b_mn_major != 0, obf = a.out_dtype == PG_BF16;
  const uint32_t osz = obf ? 2 : 4;
  const EpiArgs& e = a.epi;
  // epilogue boxes: [128 rows x (cols per chunk)] with the swizzle span of one box row; widths in ELEMENTS of each tensor
  auto emap = [&](const void* p, uint32_t esz, uint64_t width, long long ld, uint32_t box_cols, CUtensorMap* m) {
    return pg_tensor_map_2d(p, esz, width, (uint64_t)a.M, (uint64_t)ld, box_cols, BM, box_cols * esz, m);
  };
  taux = ta; tout2 = ta;                               // unused maps still have to be valid kernel parameters
  switch (a.epi_kind) {
    case EPI_STORE: case EPI_ROTARY: rc = emap(e.out, osz, a.N, e.ldo, 32, &tout); break;
    case EPI_ACCUM: rc = emap(e.out, 4, a.N, e.ldo, 32, &tout); break;
    case EPI_RESIDUAL:
      rc = emap(e.out, 4, a.N, e.ldo, 32, &tout);
      if (!rc) rc = e.aux ? emap(e.aux, 4, a.N, e.ldaux, 32, &taux) : emap(e.out, 4, a.N, e.ldo, 32, &taux);
      break;
    case EPI_GLU:
      rc = emap(e.out2, 2, a.N, e.ldo2, 32, &tout2);
      if (!rc) rc = emap(e.out, 2, a.N / 2, e.ldo, 16, &tout);
      break;
    case EPI_GELU:
      rc = emap(e.out2, 2, a.N, e.ldo2, 32, &tout2);
      if (!rc) rc = emap(e.out, 2, a.N, e.ldo, 32, &tout);
      break;
    case EPI_GLU_BWD:
      rc = emap(e.out, 2, 2ull * a.N, e.ldo, 64, &tout);
      if (!rc) rc = emap(e.aux, 2, 2ull * a.N, e.ldaux, 64, &taux);
      break;
    case EPI_GELU_BWD:
      rc = emap(e.out, 2, a.N, e.ldo, 32, &tout);
      if (!rc) rc = emap(e.aux, 2, a.N, e.ldaux, 32, &taux);
      break;
    default: break;
  }
  if (rc) return rc;
  Gemm2Dev gd{a.M, a.N, a.K, 1, a.epi};
#define TC2_CASE(BMJ, KIND, TO) return launch2_auto<false, BMJ, KIND, TO>(ta, tb, taux, tout, tout2, gd, a, stream)
  switch (a.epi_kind) {
    case EPI_ACCUM: return launch2_auto<true, true, EPI_ACCUM, float>(ta, tb, taux, tout, tout2, gd, a, stream);
    case EPI_STORE:
      if (bm) { if (obf) TC2_CASE(true, EPI_STORE, bf16); else TC2_CASE(true, EPI_STORE, float); }
      else { if (obf) TC2_CASE(false, EPI_STORE, bf16); else TC2_CASE(false, EPI_STORE, float); }
    case EPI_ROTARY: if (bm && obf) TC2_CASE(true, EPI_ROTARY, bf16); break;
    case EPI_RESIDUAL: if (bm) TC2_CASE(true, EPI_RESIDUAL, float); break;
    case EPI_GLU: if (bm && obf) TC2_CASE(true, EPI_GLU, bf16); break;
    case EPI_GELU: if (bm && obf) TC2_CASE(true, EPI_GELU, bf16); break;
    case EPI_GLU_BWD: if (!bm && obf) TC2_CASE(false, EPI_GLU_BWD, bf16); break;
    case EPI_GELU_BWD: if (!bm && obf) TC2_CASE(false, EPI_GELU_BWD, bf16); break;
    default: break;
  }
#undef TC2_CASE
  progen_set_error("gemm_tc2: unsupported combination epi=%d b_mn=%d out=%d", a.epi_kind, (int)bm, a.out_dtype);
  return PROGEN_ERR_UNSUPPORTED;
}
