// CTA-pair tcgen05 GEMM (cta_group::2): D[M,N] = A[M,K] * B[N,K]^T for the activation GEMMs (A K-major, un-batched).
//
// Why: with 128x256 tiles the K = 512 GEMMs of this model are bound by L2 -> SM operand traffic (every tile pulls
// 16 KiB of A + 32 KiB of B per k-block; measured ~10-13 TB/s aggregate).  Two CTAs of a cluster (one TPC) share a
// 256 x 256 output tile: each loads its 128 rows of A and only HALF of the B tile (16 + 16 KiB per k-block, -33 % L2
// traffic per FLOP); one tcgen05.mma.cta_group::2 (M = 256) issued by the leader CTA reads both halves and writes each
// CTA's 128 accumulator rows into that CTA's TMEM.  Everything else (TMA ring, TMEM double buffering, 8 epilogue warps
// with warp-staged IO, fused epilogues) matches gemm_tc.cu.
//
// Protocol (r = cluster rank, leader = rank 0):
//   full[s]   (leader's, count 1): leader producer arrives with expect_tx = bytes of BOTH CTAs; both CTAs' TMA loads
//             complete_tx on the leader's barrier (cp.async.bulk.tensor ... .cta_group::2 with the leader's address)
//   empty[s]  (one per CTA, count 1): tcgen05.commit.cta_group::2 ... multicast::cluster -> both producers
//   tfull[a]  (one per CTA, count 1): commit multicast -> both CTAs' epilogue warps
//   tempty[a] (leader's, count 16): 8 local + 8 remote (mapa) epilogue-warp arrivals
#include "tc_ptx.cuh"
#include "gemm.h"

namespace {

using namespace tc;
using namespace tc2;

constexpr int BM = 128;                 // rows per CTA (256 per pair)
constexpr int BN = 256;                 // columns per pair; each CTA stages BN/2 rows of B
constexpr int BK = 64;
constexpr int EW = 8;
constexpr int THREADS = 128 + 32 * EW;
constexpr int A_BYTES = BM * BK * 2;            // 16 KiB
constexpr int B_BYTES = (BN / 2) * BK * 2;      // 16 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGES = 6;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + 1024 + BAR_BYTES + EW * STAGE_WARP_BYTES;
constexpr int TMEM_COLS = 2 * BN;

struct Gemm2Dev {
  int M, N, K;
  EpiArgs epi;
};

__device__ __forceinline__ bool decode_tile2(const Gemm2Dev& g, int t, int& m0, int& n0) {
  const int m_tiles = (g.M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = g.N / BN;
  if (t >= m_tiles * n_tiles) return false;
  m0 = (t / n_tiles) * (2 * BM);
  n0 = (t % n_tiles) * BN;
  return true;
}

template <bool B_MN, int KIND, typename TO>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const Gemm2Dev g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint8_t* gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int kb_total = g.K / BK;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tma_a); prefetch_tensormap(&tma_b); }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * EW); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  cluster_sync();                                   // barriers of BOTH CTAs are initialised before anyone signals them
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - smem_base));

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int m0, n0;
      for (int t = pair; decode_tile2(g, t, m0, n0); t += npairs) {
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint32_t lead_full = mapa(full_bar(stage), 0);
          if (leader) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
          const int k0 = kb * BK;
          tma_load_2d_pair(sa, &tma_a, lead_full, k0, m0 + (int)rank * BM);                     // my 128 rows of A
          if constexpr (!B_MN) {
            tma_load_2d_pair(sb, &tma_b, lead_full, k0, n0 + (int)rank * (BN / 2));             // my 128 rows of B
          } else {
#pragma unroll
            for (int i = 0; i < BN / 2 / 64; ++i)                                               // my 2 x 64 columns of B
              tma_load_2d_pair(sb + i * 8192, &tma_b, lead_full, n0 + (int)rank * (BN / 2) + 64 * i, k0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc(2 * BM, BN, false, B_MN);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      int m0, n0;
      for (int t = pair; decode_tile2(g, t, m0, n0); t += npairs) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);              // both CTAs' epilogues drained this accumulator stage
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(full_bar(stage), phase);                    // both CTAs' operand tiles have landed
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint64_t adesc = make_smem_desc<false>(sa);
          const uint64_t bdesc = make_smem_desc<B_MN>(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_pair(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(k * (B_MN ? (2048 >> 4) : 2)), idesc,
                           (kb > 0 || k > 0) ? 1u : 0u);
          tcgen05_commit_pair(empty_bar(stage));                // frees the stage in BOTH CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tcgen05_commit_pair(tfull_bar(acc));                    // accumulators of both CTAs are complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue (each CTA drains its own 128 rows)
    const int q = warp & 3;
    const int cgroup = (warp - 4) >> 2;
    constexpr int CHUNKS_PER_GROUP = BN / (EW / 4) / 32;
    const int r_in_tile = q * 32 + lane;
    WarpStagedIO io;
    io.buf = gen + (bar_base - smem_base) + BAR_BYTES + (warp - 4) * STAGE_WARP_BYTES;
    io.lane = lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int m0, n0;
    for (int t = pair; decode_tile2(g, t, m0, n0); t += npairs) {
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const int m = m0 + (int)rank * BM + r_in_tile;
      const long long row = m;
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      const bool valid = m < g.M;
      io.valid_mask = __ballot_sync(0xffffffffu, valid);
      float rs[KIND == EPI_ROTARY ? 32 : 1], rc[KIND == EPI_ROTARY ? 32 : 1];
      bool rot_cached = false;
      if constexpr (KIND == EPI_ROTARY) {
        if (g.epi.dim_head == 64) {
          const long long pos = row % g.epi.seq_len;
          io.template load<32>(g.epi.rot_sin + pos * 32, 32, rs, true);
          io.template load<32>(g.epi.rot_cos + pos * 32, 32, rc, true);
          rot_cached = true;
        }
      }
#pragma unroll 1
      for (int c = cgroup * CHUNKS_PER_GROUP; c < (cgroup + 1) * CHUNKS_PER_GROUP; ++c) {
        const int col = n0 + c * 32;
        float v[32];
        tmem_ld32(taddr + c * 32, v);
        if constexpr (KIND == EPI_ROTARY) {
          if (rot_cached) {
            float o[32];
            if ((col & 32) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                o[i] = v[i] * rc[i >> 1] - v[i + 1] * rs[i >> 1];
                o[i + 1] = v[i + 1] * rc[i >> 1] + v[i] * rs[i >> 1];
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                o[i] = v[i] * rc[16 + (i >> 1)] - v[i + 1] * rs[16 + (i >> 1)];
                o[i + 1] = v[i + 1] * rc[16 + (i >> 1)] + v[i] * rs[16 + (i >> 1)];
              }
            }
            io.template store<32>(reinterpret_cast<TO*>(g.epi.out) + row * g.epi.ldo + col, g.epi.ldo, o, valid);
            continue;
          }
        }
        epi_apply<KIND, TO, 32>(g.epi, io, row, col, v, valid);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa(tempty_bar(acc), 0));       // the leader's barrier (local or remote)
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  cluster_sync();                                   // nobody exits (or frees TMEM) while the partner may still touch it
  if (warp == 2) { tcgen05_fence_after(); tmem_dealloc_pair<TMEM_COLS>(tmem_base); }
}

template <bool B_MN, int KIND, typename TO>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const Gemm2Dev& gd, int tiles, cudaStream_t stream) {
  auto kern = gemm_tc2_kernel<B_MN, KIND, TO>;
  static bool attr_set = false;
  if (!attr_set) {
    PG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    attr_set = true;
  }
  int pairs = pg_num_sms() / 2;
  if (tiles < pairs) pairs = tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PG_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, gd));
  PG_LAUNCH_CHECK();
  return PROGEN_OK;
}

}  // namespace

bool gemm_tc2_eligible(const GemmArgs& a) {
  static int enabled = [] { const char* e = getenv("PROGEN_GEMM_2CTA"); return e ? atoi(e) : 1; }();
  if (!enabled) return false;
  if (a.in_dtype != PG_BF16 || a.a_mn_major || a.batch != 1 || a.split_k != 1 || a.causal || a.batch_reduce) return false;
  if (a.N % BN != 0 || a.K % BK != 0 || a.M < 2 * BM) return false;
  const bool bm = a.b_mn_major != 0, obf = a.out_dtype == PG_BF16;
  switch (a.epi_kind) {                               // exactly the combinations instantiated below
    case EPI_STORE: return true;
    case EPI_ROTARY: case EPI_GLU: case EPI_GELU: return bm && obf;
    case EPI_RESIDUAL: return bm;
    case EPI_GLU_BWD: case EPI_GELU_BWD: return !bm && obf;
    default: return false;
  }
}

int gemm_tc2_launch(const GemmArgs& a, cudaStream_t stream) {
  PG_CHECK_ARG(gemm_tc2_eligible(a));
  PG_CHECK_ARG(a.lda % 8 == 0 && a.ldb % 8 == 0);
  if (a.epi_kind == EPI_ROTARY) PG_CHECK_ARG(a.epi.seq_len % 32 == 0);
  CUtensorMap ta, tb;
  int rc = pg_tensor_map_2d_bf16(a.A, a.K, a.M, a.lda, BK, BM, &ta);
  if (rc) return rc;
  if (!a.b_mn_major) rc = pg_tensor_map_2d_bf16(a.B, a.K, a.N, a.ldb, BK, BN / 2, &tb);
  else               rc = pg_tensor_map_2d_bf16(a.B, a.N, a.K, a.ldb, 64, BK, &tb);
  if (rc) return rc;
  Gemm2Dev gd{a.M, a.N, a.K, a.epi};
  const int tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * (a.N / BN);
  const bool bm = a.b_mn_major != 0, obf = a.out_dtype == PG_BF16;
#define TC2_CASE(BMJ, KIND, TO) return launch2<BMJ, KIND, TO>(ta, tb, gd, tiles, stream)
  switch (a.epi_kind) {
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
