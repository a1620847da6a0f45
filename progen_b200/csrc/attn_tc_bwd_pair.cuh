// Paired forms of the two tcgen05 attention-backward kernels (included by attn_tc_bwd.cu inside its anonymous
// namespace; window % 256 == 0).  Same math, tiles, shared-memory layouts and MMA shapes as attn_bwd_dq_tc_kernel /
// attn_bwd_dkv_tc_kernel; what changes is WHO works on WHAT:
//
//   * a CTA owns a PAIR of adjacent 128-row work items (A, B) of one window and streams the column tiles they share
//     ONCE (K_j|V_j for dQ, Q_j|dO_j for dK/dV): half the TMA traffic of two single items;
//   * each item has its own group of 4 element-wise warps with thread == row (TMEM lane == row), so nothing is
//     exchanged between threads (the single-item kernels split a row between two warps and meet in a 256-thread
//     barrier every step) and the two groups drift freely: while one group waits for its MMAs the other computes;
//   * each group has its OWN MMA-issuing thread, which issues the group's NEXT S / dP as soon as the group has read
//     the current one out of TMEM, i.e. before its dS tile is written, so the score MMAs run under the element-wise
//     work of the same group and neither group's ready work ever queues behind the other group's barrier.
// The single-item kernels stay for window % 256 != 0.

// ===================================================================================================== dQ, paired
namespace dqp {
constexpr int KV_STAGES = 4;
constexpr int KV_BYTES = 2 * COL_TILE_BYTES;                                   // K_j then V_j
constexpr int OFF_QDO = 0;                                                     // [g]: Q_g, dO_g
constexpr int OFF_KV = 4 * ROW_TILE_BYTES;
constexpr int OFF_DS = OFF_KV + KV_STAGES * KV_BYTES;                          // [g][2 bufs] dS tiles
constexpr int OFF_BAR = OFF_DS + 4 * ES_BYTES;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + 1024;
constexpr int TM_GROUP = 192;                                                  // TMEM columns per group: S 64 | dP 64 | dQ 64
}  // namespace dqp

struct QPair { int b, hh, q0, win, i0, nprev, nA; };
__device__ __forceinline__ bool decode_qpair(const BwdDev& a, int wi, QPair& it) {
  const int pairs = a.n / (2 * RB);
  if (wi >= a.B * a.h * pairs) return false;
  const int p = wi % pairs, r = wi / pairs;
  it.hh = r % a.h; it.b = r / a.h;
  it.q0 = p * 2 * RB; it.win = it.q0 / a.w; it.i0 = it.q0 % a.w;
  it.nprev = it.win > 0 ? a.w / CT : 0;                  // zero look-back keys of window 0 carry no gradient (K == 0)
  it.nA = it.nprev + (it.i0 + RB) / CT;                  // 64-key tiles seen by A; B sees nA + 2
  return true;
}

__global__ void __launch_bounds__(384, 1) attn_bwd_dq_pair_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                 const __grid_constant__ CUtensorMap tmap_kv,
                                                                 const __grid_constant__ CUtensorMap tmap_do, const BwdDev a) {
  using namespace dqp;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQDO = base + OFF_QDO, sKV = base + OFF_KV, sDS = base + OFF_DS, bars = base + OFF_BAR;
  const uint32_t qdo_full = bars, qdo_empty = bars + 8;
  auto kv_full = [&](int s) { return bars + 16 + 8 * s; };
  auto kv_empty = [&](int s) { return bars + 48 + 8 * s; };
  auto sd_full = [&](int g) { return bars + 80 + 8 * g; };
  auto sd_empty = [&](int g) { return bars + 96 + 8 * g; };
  auto ds_full = [&](int g, int b) { return bars + 112 + 8 * (2 * g + b); };
  auto ds_empty = [&](int g, int b) { return bars + 144 + 8 * (2 * g + b); };
  auto dq_full = [&](int g) { return bars + 176 + 8 * g; };
  auto dq_empty = [&](int g) { return bars + 192 + 8 * g; };
  const uint32_t tmem_slot = bars + 208;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = a.h * DH;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmap_qkv); prefetch_tensormap(&tmap_kv); prefetch_tensormap(&tmap_do); }
  if (warp == 1 && lane == 0) {
    mbar_init(qdo_full, 1); mbar_init(qdo_empty, 2);                            // one release per group's MMA issuer
    for (int s = 0; s < KV_STAGES; ++s) { mbar_init(kv_full(s), 1); mbar_init(kv_empty(s), 2); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(sd_full(g), 1); mbar_init(sd_empty(g), 4); mbar_init(dq_full(g), 1); mbar_init(dq_empty(g), 4);
      for (int b = 0; b < 2; ++b) { mbar_init(ds_full(g, b), 4); mbar_init(ds_empty(g, b), 1); }
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + OFF_BAR + 208);
  auto key_pos = [&](const QPair& it, int kt) { return kt < it.nprev ? (it.win - 1) * a.w + kt * CT : it.win * a.w + (kt - it.nprev) * CT; };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t kv_phase = 0, q_phase = 0;
      QPair it;
      for (int wi = blockIdx.x; decode_qpair(a, wi, it); wi += gridDim.x) {
        const int row0 = it.b * a.n;
        mbar_wait(qdo_empty, q_phase ^ 1);
        mbar_expect_tx(qdo_full, 4 * ROW_TILE_BYTES);
        for (int g = 0; g < 2; ++g) {
          tma_load_2d(sQDO + (2 * g) * ROW_TILE_BYTES, &tmap_qkv, qdo_full, it.hh * DH, row0 + it.q0 + g * RB);
          tma_load_2d(sQDO + (2 * g + 1) * ROW_TILE_BYTES, &tmap_do, qdo_full, it.hh * DH, row0 + it.q0 + g * RB);
        }
        q_phase ^= 1;
        for (int kt = 0; kt < it.nA + 2; ++kt) {
          mbar_wait(kv_empty(stage), kv_phase ^ 1);
          const uint32_t dst = sKV + stage * KV_BYTES;
          const int kp = row0 + key_pos(it, kt);
          mbar_expect_tx(kv_full(stage), KV_BYTES);
          tma_load_2d(dst, &tmap_kv, kv_full(stage), I + it.hh * DH, kp);                       // 64-row boxes
          tma_load_2d(dst + COL_TILE_BYTES, &tmap_kv, kv_full(stage), 2 * I + it.hh * DH, kp);
          if (++stage == KV_STAGES) { stage = 0; kv_phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // one MMA issuer PER GROUP (warp 1 -> A, warp 3 -> B): a single in-order issuer serving both groups blocks one
    // group's ready work behind the other group's barrier (ncu: 21 % of the samples in the S/dP wait)
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(RB, CT, false, false);      // [128 x 64] = A (K-major) x B^T (K-major), K = dh
      constexpr uint32_t idesc_a = make_idesc(RB, DH, false, true);       // [128 x 64] += dS (K-major, K = keys) x K_j (MN-major)
      const int g = warp == 1 ? 0 : 1;
      int stage = 0;
      uint32_t kv_phase = 0, q_phase = 0, item = 0;
      uint32_t s_issued = 0, ds_used = 0;
      QPair it;
      const uint32_t tm = tmem_base + g * TM_GROUP;
      auto issue_s = [&](int st) {
        if (s_issued > 0) mbar_wait(sd_empty(g), (s_issued - 1) & 1);             // the group has read its previous S / dP
        ++s_issued;
        tcgen05_fence_after();
        const uint64_t qd = make_smem_desc<false>(sQDO + (2 * g) * ROW_TILE_BYTES);
        const uint64_t dod = make_smem_desc<false>(sQDO + (2 * g + 1) * ROW_TILE_BYTES);
        const uint64_t kd = make_smem_desc<false>(sKV + st * KV_BYTES), vd = make_smem_desc<false>(sKV + st * KV_BYTES + COL_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tm, qd + 2 * k, kd + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tm + 64, dod + 2 * k, vd + 2 * k, idesc_s, k > 0);
        tcgen05_commit(sd_full(g));
      };
      for (int wi = blockIdx.x; decode_qpair(a, wi, it); wi += gridDim.x, ++item) {
        mbar_wait(qdo_full, q_phase);
        q_phase ^= 1;
        const int n_g = it.nA + 2 * g, n_all = it.nA + 2;
        mbar_wait(kv_full(stage), kv_phase);
        issue_s(stage);
        for (int j = 0; j < n_all; ++j) {
          int nstage = stage + 1;
          uint32_t nphase = kv_phase;
          if (nstage == KV_STAGES) { nstage = 0; nphase ^= 1; }
          if (j < n_g) {
            if (j + 1 < n_g) {
              mbar_wait(kv_full(nstage), nphase);
              issue_s(nstage);
            } else {
              tcgen05_commit(qdo_empty);                                   // this group's Q / dO tiles are no longer needed
            }
            const uint32_t buf = ds_used & 1;
            mbar_wait(ds_full(g, buf), (ds_used >> 1) & 1);
            ++ds_used;
            if (j == 0 && item > 0) mbar_wait(dq_empty(g), (item - 1) & 1);   // previous item's dQ has been read out
            tcgen05_fence_after();
            const uint64_t dsd = make_smem_desc<false>(sDS + (2 * g + buf) * ES_BYTES);
            const uint64_t kmn = make_smem_desc<true>(sKV + stage * KV_BYTES);
#pragma unroll
            for (int k = 0; k < CT / 16; ++k)
              umma_bf16(tm + 128, dsd + 2 * k, kmn + (uint64_t)(k * (2048 >> 4)), idesc_a, (j > 0 || k > 0) ? 1u : 0u);
            tcgen05_commit(ds_empty(g, buf));
            if (j == n_g - 1) tcgen05_commit(dq_full(g));
            tcgen05_commit(kv_empty(stage));                               // my MMAs on K_j / V_j (the other issuer adds its own)
          } else {
            // a tile only the other group uses: wait until it has landed (so the arrival lands in the right phase)
            mbar_wait(kv_full(stage), kv_phase);
            mbar_arrive(kv_empty(stage));
          }
          stage = nstage;
          kv_phase = nphase;
        }
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, g = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tm = tmem_base + g * TM_GROUP + lane_addr;
    const float scale = 0.125f, sc = 0.125f * LOG2E;
    uint32_t cnt = 0, item = 0;                                               // tiles processed by this group; items
    QPair it, nx;
    // per-row constants of an item: delta = rowsum(dO o O) over the 64 channels of this row (also written out for the
    // dK/dV kernel) and lse in log2 units.  Computed for the NEXT item while the MMAs finish the current item's dQ, so
    // the global-load latency of the prologue is not exposed at every item start (13 % of the samples before).
    auto row_consts = [&](const QPair& p, float& D, float& L2) {
      const long long t = (long long)p.b * a.n + p.q0 + g * RB + row;
      D = 0.f;
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        float o[32], d[32];
        load_vec<32>(a.out + t * I + p.hh * DH + hc * 32, o);
        load_vec<32>(a.dout + t * I + p.hh * DH + hc * 32, d);
#pragma unroll
        for (int i = 0; i < 32; ++i) D = fmaf(o[i], d[i], D);
      }
      a.delta[t * a.h + p.hh] = D;
      L2 = a.lse[t * a.h + p.hh] * LOG2E;
    };
    float D = 0.f, L2 = 0.f;
    if (decode_qpair(a, blockIdx.x, it)) row_consts(it, D, L2);
    for (int wi = blockIdx.x; decode_qpair(a, wi, it); wi += gridDim.x, ++item) {
      const long long t = (long long)it.b * a.n + it.q0 + g * RB + row;
      const int qi = it.i0 + g * RB + row;
      const int nt = it.nA + 2 * g;
      for (int j = 0; j < nt; ++j, ++cnt) {
        const uint32_t buf = cnt & 1;
        const bool own = j >= it.nprev;
        const int c0 = (j - it.nprev) * CT;                                   // in-window offset of the tile's first key
        mbar_wait(sd_full(g), cnt & 1);
        tcgen05_fence_after();
        if (cnt >= 2) mbar_wait(ds_empty(g, buf), ((cnt - 2) >> 1) & 1);      // MMA finished reading the tile of step cnt-2
        uint8_t* tile = gen + OFF_DS + (2 * g + buf) * ES_BYTES;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float s[32], dp[32];
          tmem_ld32x2(tm + half * 32, tm + 64 + half * 32, s, dp);
          if (half == 1) {                                                    // S / dP are in registers: the next pair may be issued
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sd_empty(g));
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float p = ex2_approx(s[i] * sc - L2);
            if (own && c0 + half * 32 + i > qi) p = 0.f;
            s[i] = p * (dp[i] - D) * scale;
          }
          write_es_row(tile, row, half, s);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(ds_full(g, buf));
      }
      float Dn = 0.f, L2n = 0.f;
      if (decode_qpair(a, wi + gridDim.x, nx)) row_consts(nx, Dn, L2n);
      mbar_wait(dq_full(g), item & 1);
      tcgen05_fence_after();
      float dqa[32], dqb[32];
      tmem_ld32x2(tm + 128, tm + 160, dqa, dqb);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty(g));
      bf16* dst = a.dqkv + t * (3LL * I) + it.hh * DH;
      store_grad_row(a, dst, it.q0 + g * RB + row, 0, dqa);
      store_grad_row(a, dst + 32, it.q0 + g * RB + row, 32, dqb);
      D = Dn; L2 = L2n;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) { tcgen05_fence_after(); tmem_dealloc<TMEM_COLS>(tmem_base); }
}

// ===================================================================================================== dK, dV, paired
namespace dkvp {
constexpr int Q_STAGES = 3;
constexpr int QS_BYTES = 2 * COL_TILE_BYTES;                                   // Q_j then dO_j
constexpr int OFF_KV = 0;                                                      // [g]: K_g, V_g
constexpr int OFF_QS = 4 * ROW_TILE_BYTES;
constexpr int OFF_ES = OFF_QS + Q_STAGES * QS_BYTES;                           // [g]: P^T | dS^T
constexpr int OFF_STAT = OFF_ES + 4 * ES_BYTES;                                // [stage][lse*log2e | delta][64]
constexpr int STAT_BYTES = 2 * CT * 4;
constexpr int OFF_BAR = OFF_STAT + Q_STAGES * STAT_BYTES;
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = OFF_BAR + BAR_BYTES + 1024;
constexpr int TM_GROUP = 256;                                                  // S^T 64 | dP^T 64 | dK 64 | dV 64
}  // namespace dkvp

struct KPair { int b, hh, k0, win, j0, nown, ntiles; };
__device__ __forceinline__ bool decode_kpair(const BwdDev& a, int wi, KPair& it) {
  const int pairs = a.n / (2 * RB);
  if (wi >= a.B * a.h * pairs) return false;
  const int p = wi % pairs, r = wi / pairs;
  it.hh = r % a.h; it.b = r / a.h;
  it.k0 = p * 2 * RB; it.win = it.k0 / a.w; it.j0 = it.k0 % a.w;
  it.nown = (a.w - it.j0) / CT;                                                // A's own-window query tiles from its diagonal on
  it.ntiles = it.nown + ((it.win + 1 < a.n / a.w) ? a.w / CT : 0);            // + the whole next window
  return true;                                                                 // B (keys k0+128..) skips the first two tiles
}

__global__ void __launch_bounds__(384, 1) attn_bwd_dkv_pair_kernel(const __grid_constant__ CUtensorMap tmap_qkv_row,
                                                                  const __grid_constant__ CUtensorMap tmap_qkv_col,
                                                                  const __grid_constant__ CUtensorMap tmap_do_col, const BwdDev a) {
  using namespace dkvp;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sKVr = base + OFF_KV, sQS = base + OFF_QS, sES = base + OFF_ES, bars = base + OFF_BAR;
  const uint32_t kvi_full = bars, kvi_empty = bars + 8;
  auto qs_full = [&](int s) { return bars + 16 + 8 * s; };
  auto qs_empty = [&](int s) { return bars + 40 + 8 * s; };
  auto st_full = [&](int g) { return bars + 64 + 8 * g; };
  auto st_empty = [&](int g) { return bars + 80 + 8 * g; };
  auto es_full = [&](int g) { return bars + 96 + 8 * g; };
  auto es_empty = [&](int g) { return bars + 112 + 8 * g; };
  auto acc_full = [&](int g) { return bars + 128 + 8 * g; };
  auto acc_empty = [&](int g) { return bars + 144 + 8 * g; };
  const uint32_t tmem_slot = bars + 160;
  float* stats = reinterpret_cast<float*>(gen + OFF_STAT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int I = a.h * DH;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmap_qkv_row); prefetch_tensormap(&tmap_qkv_col); prefetch_tensormap(&tmap_do_col); }
  if (warp == 1 && lane == 0) {
    mbar_init(kvi_full, 1); mbar_init(kvi_empty, 2);                            // one release per group's MMA issuer
    for (int s = 0; s < Q_STAGES; ++s) { mbar_init(qs_full(s), 2); mbar_init(qs_empty(s), 2); }   // full: TMA bytes + the stats warp
    for (int g = 0; g < 2; ++g) {
      mbar_init(st_full(g), 1); mbar_init(st_empty(g), 4); mbar_init(es_full(g), 4); mbar_init(es_empty(g), 1);
      mbar_init(acc_full(g), 1); mbar_init(acc_empty(g), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + OFF_BAR + 160);
  auto q_pos = [&](const KPair& it, int qt) { return qt < it.nown ? it.win * a.w + it.j0 + qt * CT : (it.win + 1) * a.w + (qt - it.nown) * CT; };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t q_phase = 0, kv_phase = 0;
      KPair it;
      for (int wi = blockIdx.x; decode_kpair(a, wi, it); wi += gridDim.x) {
        const int row0 = it.b * a.n;
        mbar_wait(kvi_empty, kv_phase ^ 1);
        mbar_expect_tx(kvi_full, 4 * ROW_TILE_BYTES);
        for (int g = 0; g < 2; ++g) {
          tma_load_2d(sKVr + (2 * g) * ROW_TILE_BYTES, &tmap_qkv_row, kvi_full, I + it.hh * DH, row0 + it.k0 + g * RB);
          tma_load_2d(sKVr + (2 * g + 1) * ROW_TILE_BYTES, &tmap_qkv_row, kvi_full, 2 * I + it.hh * DH, row0 + it.k0 + g * RB);
        }
        kv_phase ^= 1;
        for (int qt = 0; qt < it.ntiles; ++qt) {
          mbar_wait(qs_empty(stage), q_phase ^ 1);
          const uint32_t dst = sQS + stage * QS_BYTES;
          const int qp = row0 + q_pos(it, qt);
          mbar_expect_tx(qs_full(stage), QS_BYTES);
          tma_load_2d(dst, &tmap_qkv_col, qs_full(stage), it.hh * DH, qp);
          tma_load_2d(dst + COL_TILE_BYTES, &tmap_do_col, qs_full(stage), it.hh * DH, qp);
          if (++stage == Q_STAGES) { stage = 0; q_phase ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // per-column constants of every query tile (lse in log2 units, delta): plain loads, published with the stage
    int stage = 0;
    uint32_t q_phase = 0;
    KPair it;
    for (int wi = blockIdx.x; decode_kpair(a, wi, it); wi += gridDim.x) {
      const long long row0 = (long long)it.b * a.n;
      for (int qt = 0; qt < it.ntiles; ++qt) {
        mbar_wait(qs_empty(stage), q_phase ^ 1);
        float* xl = stats + stage * (2 * CT);
        const long long qp = row0 + q_pos(it, qt);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const long long idx = (qp + r * 32 + lane) * a.h + it.hh;
          xl[r * 32 + lane] = a.lse[idx] * LOG2E;
          xl[CT + r * 32 + lane] = a.delta[idx];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(qs_full(stage));
        if (++stage == Q_STAGES) { stage = 0; q_phase ^= 1; }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // one MMA issuer per group (warp 1 -> A, warp 2 -> B once it has allocated TMEM)
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(RB, CT, false, false);      // S^T / dP^T [128 keys x 64 queries], K = dh
      constexpr uint32_t idesc_a = make_idesc(RB, DH, false, true);       // dV / dK [128 keys x 64 dh], K = queries, B MN-major
      const int g = warp == 1 ? 0 : 1;
      const int first_t = 2 * g;                                          // B's keys are not visible to the first two query tiles
      int stage = 0;
      uint32_t q_phase = 0, kv_phase = 0, item = 0;
      uint32_t st_issued = 0, es_used = 0;
      KPair it;
      const uint32_t tm = tmem_base + g * TM_GROUP;
      auto issue_st = [&](int st) {
        if (st_issued > 0) mbar_wait(st_empty(g), (st_issued - 1) & 1);           // the group has read its previous S^T / dP^T
        ++st_issued;
        tcgen05_fence_after();
        const uint64_t kd = make_smem_desc<false>(sKVr + (2 * g) * ROW_TILE_BYTES);
        const uint64_t vd = make_smem_desc<false>(sKVr + (2 * g + 1) * ROW_TILE_BYTES);
        const uint64_t qd = make_smem_desc<false>(sQS + st * QS_BYTES), dod = make_smem_desc<false>(sQS + st * QS_BYTES + COL_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tm, kd + 2 * k, qd + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) umma_bf16(tm + 64, vd + 2 * k, dod + 2 * k, idesc_s, k > 0);
        tcgen05_commit(st_full(g));
      };
      for (int wi = blockIdx.x; decode_kpair(a, wi, it); wi += gridDim.x, ++item) {
        mbar_wait(kvi_full, kv_phase);
        kv_phase ^= 1;
        const int nT = it.ntiles;
        for (int t = 0; t < nT; ++t) {
          int nstage = stage + 1;
          uint32_t nphase = q_phase;
          if (nstage == Q_STAGES) { nstage = 0; nphase ^= 1; }
          if (t < first_t) {
            // a tile only the other group uses: wait until it has landed (so the arrival lands in the right phase)
            mbar_wait(qs_full(stage), q_phase);
            mbar_arrive(qs_empty(stage));
          } else {
            if (t == first_t) {
              mbar_wait(qs_full(stage), q_phase);
              issue_st(stage);
            }
            if (t + 1 < nT) {
              mbar_wait(qs_full(nstage), nphase);
              issue_st(nstage);
            } else {
              tcgen05_commit(kvi_empty);                                   // this group's K / V row tiles are free for the next item
            }
            const bool first = t == first_t;
            mbar_wait(es_full(g), es_used & 1);
            ++es_used;
            if (first && item > 0) mbar_wait(acc_empty(g), (item - 1) & 1);
            tcgen05_fence_after();
            const uint64_t ptd = make_smem_desc<false>(sES + (2 * g) * ES_BYTES);
            const uint64_t dsd = make_smem_desc<false>(sES + (2 * g + 1) * ES_BYTES);
            const uint64_t qmn = make_smem_desc<true>(sQS + stage * QS_BYTES);
            const uint64_t domn = make_smem_desc<true>(sQS + stage * QS_BYTES + COL_TILE_BYTES);
#pragma unroll
            for (int k = 0; k < CT / 16; ++k)                              // dV += P^T dO_j
              umma_bf16(tm + 192, ptd + 2 * k, domn + (uint64_t)(k * (2048 >> 4)), idesc_a, (!first || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < CT / 16; ++k)                              // dK += dS^T Q_j
              umma_bf16(tm + 128, dsd + 2 * k, qmn + (uint64_t)(k * (2048 >> 4)), idesc_a, (!first || k > 0) ? 1u : 0u);
            tcgen05_commit(es_empty(g));
            tcgen05_commit(qs_empty(stage));                               // my MMAs on Q_j / dO_j (the other issuer adds its own)
          }
          stage = nstage;
          q_phase = nphase;
        }
        tcgen05_commit(acc_full(g));
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3, g = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t tm = tmem_base + g * TM_GROUP + lane_addr;
    const float scale = 0.125f, sc = 0.125f * LOG2E;
    int stage = 0;
    uint32_t q_phase = 0, cnt = 0, item = 0;
    KPair it;
    for (int wi = blockIdx.x; decode_kpair(a, wi, it); wi += gridDim.x, ++item) {
      const long long row0 = (long long)it.b * a.n;
      const int kj = it.j0 + g * RB + row;                                     // in-window offset of this thread's key row
      for (int t = 0; t < it.ntiles; ++t) {
        const int st = stage;
        const uint32_t ph = q_phase;
        if (++stage == Q_STAGES) { stage = 0; q_phase ^= 1; }
        if (g == 1 && t < 2) continue;                                         // queries before B's keys: nothing to do
        const bool own = t < it.nown;
        const int c0 = it.j0 + t * CT;                                         // in-window offset of the tile's first query
        mbar_wait(qs_full(st), ph);                                            // the stage's lse / delta columns are visible
        const float* xl = stats + st * (2 * CT);
        mbar_wait(st_full(g), cnt & 1);
        tcgen05_fence_after();
        if (cnt >= 1) mbar_wait(es_empty(g), (cnt - 1) & 1);                   // previous P^T / dS^T consumed by the MMAs
        uint8_t* pt = gen + OFF_ES + (2 * g) * ES_BYTES;
        uint8_t* dst = pt + ES_BYTES;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float s[32], dp[32];
          tmem_ld32x2(tm + half * 32, tm + 64 + half * 32, s, dp);
          if (half == 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(st_empty(g));
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float p = ex2_approx(s[i] * sc - xl[half * 32 + i]);
            if (own && kj > c0 + half * 32 + i) p = 0.f;                       // key after query: masked
            dp[i] = p * (dp[i] - xl[CT + half * 32 + i]) * scale;              // dS^T
            s[i] = p;                                                          // P^T
          }
          write_es_row(pt, row, half, s);
          write_es_row(dst, row, half, dp);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(es_full(g));
        ++cnt;
      }
      mbar_wait(acc_full(g), item & 1);
      tcgen05_fence_after();
      const long long tr = row0 + it.k0 + g * RB + row;
      bf16* out = a.dqkv + tr * (3LL * I) + I + it.hh * DH;
      {
        float x0[32], x1[32];
        tmem_ld32x2(tm + 128, tm + 160, x0, x1);                               // dK
        store_grad_row(a, out, it.k0 + g * RB + row, 0, x0);
        store_grad_row(a, out + 32, it.k0 + g * RB + row, 32, x1);
      }
      {
        float x0[32], x1[32];
        tmem_ld32x2(tm + 192, tm + 224, x0, x1);                               // dV
        store_grad_row(a, out + I, it.k0 + g * RB + row, 0, x0);
        store_grad_row(a, out + I + 32, it.k0 + g * RB + row, 32, x1);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(g));
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) { tcgen05_fence_after(); tmem_dealloc<TMEM_COLS>(tmem_base); }
}
