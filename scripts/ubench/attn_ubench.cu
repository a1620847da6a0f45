// Micro-benchmarks behind the attention kernel design (sm_100a): what one SM really does per clock for the pieces of a
// softmax / score step.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I progen_b200/csrc -I include
//   scripts/ubench/attn_ubench.cu -o gpurun_out/attn_ubench -lcuda ; run on a B200.  One CTA per test (grid 1) unless stated.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_ptx.cuh"

using namespace tc;

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

// ---------------------------------------------------------------- 1. XU / FMA / ALU issue rates with W warps per sub-partition
template <int KIND>
__global__ void pipe_rate_kernel(float* out, long long* cyc, int iters) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = -(float)(threadIdx.x + i) * 1e-3f;
  float2 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = make_float2(v[2 * i], v[2 * i + 1]);
  __syncthreads();
  const long long t0 = clk();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (KIND == 0) v[i] = ex2f(v[i]);                                        // MUFU.EX2
      if (KIND == 1) v[i] = fmaf(v[i], 1.0001f, 0.5f);                         // FFMA
      if (KIND == 3) v[i] = fmaxf(v[i], v[(i + 1) & 15] * 0.5f);               // FMNMX (+FMUL)
      if (KIND == 4) v[i] = __uint_as_float(pack_bf16x2(v[i], v[(i + 3) & 15]));   // F2FP
    }
    if (KIND == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = ffma2(p[i], make_float2(1.0001f, 1.0001f), make_float2(0.5f, 0.5f));   // FFMA2
    }
  }
  const long long t1 = clk();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += p[i].x + p[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// ---------------------------------------------------------------- 2. the softmax exp phase of one 128-column row (as in attn_fwd_ts)
template <bool POLY>
__global__ void exp_phase_kernel(float* out, long long* cyc, int iters) {
  uint32_t s[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) s[c][i] = __float_as_uint(-(float)((threadIdx.x * 7 + c * 32 + i) % 97) * 0.1f);
  const float sc = 0.18f, m_used = 0.3f;
  float acc = 0.f;
  uint32_t keep = 0;
  __syncthreads();
  const long long t0 = clk();
  for (int it = 0; it < iters; ++it) {
    const float2 sc2 = make_float2(sc, sc), nm2 = make_float2(-m_used - it * 1e-6f, -m_used);
    float2 rs0 = make_float2(0.f, 0.f), rs1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float2 x0 = ffma2(make_float2(__uint_as_float(s[c][i]), __uint_as_float(s[c][i + 1])), sc2, nm2);
        float2 x1 = ffma2(make_float2(__uint_as_float(s[c][i + 2]), __uint_as_float(s[c][i + 3])), sc2, nm2);
        x0.x = ex2f(x0.x); x0.y = ex2f(x0.y); x1.x = ex2f(x1.x);
        x1.y = POLY ? ex2_poly(x1.y) : ex2f(x1.y);
        rs0 = fadd2(rs0, x0); rs1 = fadd2(rs1, x1);
        keep ^= pack_bf16x2(x0.x, x0.y) + pack_bf16x2(x1.x, x1.y);
      }
    }
    acc += rs0.x + rs0.y + rs1.x + rs1.y;
  }
  const long long t1 = clk();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float(keep & 0xffff);
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// ---------------------------------------------------------------- 3. TMEM load / store: latency of 4 x (32x32b.x32) + wait, W warps at once
__global__ void tmem_ldst_kernel(float* out, long long* cyc, int iters, int do_store) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t r[4][32];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) r[c][i] = threadIdx.x + i;
  for (int c = 0; c < 4; ++c) tmem_st<32>(base + c * 32, r[c]);
  tmem_st_wait();
  __syncthreads();
  const long long t0 = clk();
  for (int it = 0; it < iters; ++it) {
    if (do_store) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st<32>(base + c * 32, r[c]);
      tmem_st_wait();
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32_issue(base + c * 32, r[c]);
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32_wait(r[c]);
    }
  }
  const long long t1 = clk();
  uint32_t x = 0;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= r[c][i];
  out[threadIdx.x] = __uint_as_float(x);
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(slot); }
}

// ---------------------------------------------------------------- 4. MMA: issue n groups back to back, commit, wait -> cycles
// kind 0: SS 128x128x64 (4 x K16)   kind 1: SS 128x64x64 (4 x K16)   kind 2: TS 128x64x128 (8 x K16, A from TMEM)   kind 3: TS 128x64x64
// concurrent_ld: warps 4..7 hammer tcgen05.ld meanwhile (TMEM port contention)
__global__ void mma_kernel(long long* cyc, int kind, int ngroups, int concurrent_ld, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  // operands: zeros are fine for timing
  for (int i = threadIdx.x; i < 64 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    const uint64_t ad = make_smem_desc<false>(base), bd = make_smem_desc<false>(base + 16384), bmn = make_smem_desc<true>(base + 32768);
    const uint32_t id_s128 = make_idesc(128, 128, false, false), id_s64 = make_idesc(128, 64, false, false), id_a = make_idesc(128, 64, false, true);
    const long long t0 = clk();
    for (int g = 0; g < ngroups; ++g) {
      if (kind == 0) for (int k = 0; k < 4; ++k) umma_bf16(tm, ad + 2 * k, bd + 2 * k, id_s128, k > 0);
      if (kind == 1) for (int k = 0; k < 4; ++k) umma_bf16(tm, ad + 2 * k, bd + 2 * k, id_s64, k > 0);
      if (kind == 2) for (int k = 0; k < 8; ++k) umma_bf16_ts(tm + 256, tm + 384 + 8 * k, bmn + (uint64_t)(k * (2048 >> 4)), id_a, 1);
      if (kind == 3) for (int k = 0; k < 4; ++k) umma_bf16_ts(tm + 256, tm + 384 + 8 * k, bmn + (uint64_t)(k * (2048 >> 4)), id_a, 1);
    }
    const long long t1 = clk();
    tcgen05_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clk();
    cyc[0] = t1 - t0;
    cyc[1] = t2 - t0;
    stop = 1;
  } else if (warp >= 4 && concurrent_ld) {
    const uint32_t a = tm + ((uint32_t)((warp & 3) * 32) << 16) + 128;
    uint32_t r[32];
    uint32_t x = 0;
    long long n = 0;
    while (!stop) {
      tmem_ld32_issue(a, r);
      tmem_ld32_wait(r);
      x ^= r[lane];
      ++n;
    }
    out[threadIdx.x] = __uint_as_float(x);
    if (threadIdx.x == 128) cyc[2] = n;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(slot); }
}


// ---------------------------------------------------------------- 4b. MMA cost table: 32 back-to-back K=16 instructions, M = 128, N and operand source vary
__global__ void mma_sweep_kernel(long long* cyc, int N, int ts, int b_mn, int nacc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    const uint64_t ad = make_smem_desc<false>(base);
    const uint64_t bk = make_smem_desc<false>(base + 16384), bm = make_smem_desc<true>(base + 16384, 16384);
    const uint32_t idesc = make_idesc(128, N, false, b_mn != 0);
    // warm-up
    for (int k = 0; k < 4; ++k) umma_bf16(tm, ad + 2 * k, bk + 2 * k, make_idesc(128, 64, false, false), k > 0);
    tcgen05_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t0 = clk();
    for (int i = 0; i < 32; ++i) {
      const int k = i & 3;
      const uint32_t d = tm + (nacc > 1 ? (i >> 2) % nacc * N : 0);          // nacc accumulators in rotation (independent chains)
      const uint64_t bd = b_mn ? bm + (uint64_t)(k * (2048 >> 4)) : bk + 2 * k;
      if (ts) umma_bf16_ts(d, tm + 448 + 8 * k, bd, idesc, 1);
      else umma_bf16(d, ad + 2 * k, bd, idesc, 1);
    }
    tcgen05_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 1);
    const long long t2 = clk();
    cyc[0] = t2 - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(slot); }
}


__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
// 4c. same sweep, but the WHOLE warp runs the issue loop (descriptors are warp-uniform values) and only the MMA itself is
// predicated on an elected lane: does ptxas keep the operands in uniform registers instead of the R2UR waterfall?
__global__ void mma_sweep_uniform_kernel(long long* cyc, int N, int ts, int mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint64_t ad = make_smem_desc<false>(base);
    const uint64_t bk = make_smem_desc<false>(base + 16384);
    const uint32_t idesc = make_idesc(128, N, false, false);
    const long long t0 = clk();
    if (mode == 0) {                      // elect once, loop inside (round-1 style but with uniform inputs)
      if (elect_one()) {
#pragma unroll 1
        for (int g = 0; g < 8; ++g) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ts) umma_bf16_ts(tm, tm + 448 + 8 * k, bk + 2 * k, idesc, 1);
            else umma_bf16(tm, ad + 2 * k, bk + 2 * k, idesc, 1);
          }
        }
      }
    } else {                              // whole warp loops, elect per group
#pragma unroll 1
      for (int g = 0; g < 8; ++g) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ts) umma_bf16_ts(tm, tm + 448 + 8 * k, bk + 2 * k, idesc, 1);
            else umma_bf16(tm, ad + 2 * k, bk + 2 * k, idesc, 1);
          }
        }
        __syncwarp();
      }
    }
    const long long t1 = clk();
    if (elect_one()) tcgen05_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clk();
    if ((threadIdx.x & 31) == 0) { cyc[0] = t2 - t0; cyc[1] = t1 - t0; }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(slot); }
}

// ---------------------------------------------------------------- 5. mbarrier ping-pong between two warps (round trips)
__global__ void pingpong_kernel(long long* cyc, int iters) {
  __shared__ __align__(8) uint64_t b0, b1;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&b0), 1); mbar_init(smem_u32(&b1), 1); fence_barrier_init(); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long t0 = clk();
  for (int it = 0; it < iters; ++it) {
    if (warp == 0) {
      if (lane == 0) mbar_arrive(smem_u32(&b0));
      mbar_wait(smem_u32(&b1), it & 1);
    } else {
      mbar_wait(smem_u32(&b0), it & 1);
      if (lane == 0) mbar_arrive(smem_u32(&b1));
    }
  }
  const long long t1 = clk();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  long long h[8];
  auto get = [&]() { cudaDeviceSynchronize(); cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); } cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost); };
  const int IT = 2000;
  const char* names[] = {"MUFU.EX2", "FFMA", "FFMA2", "FMNMX+FMUL", "F2FP"};
  for (int warps_per_smsp : {1, 2, 4}) {
    const int threads = 128 * warps_per_smsp;
#define RATE(K) pipe_rate_kernel<K><<<1, threads>>>(out, cyc, IT); get(); \
    printf("pipe %-11s warps/SMSP %d: %.2f cycles per warp-instruction per SMSP\n", names[K], warps_per_smsp, (double)h[0] / (IT * (K == 2 ? 8 : 16) * warps_per_smsp));
    RATE(0) RATE(1) RATE(2) RATE(3) RATE(4)
  }
  for (int warps_per_smsp : {1, 2}) {
    exp_phase_kernel<false><<<1, 128 * warps_per_smsp>>>(out, cyc, 200); get();
    printf("exp phase (128 elements/thread, MUFU only)  warps/SMSP %d: %.0f cycles per warp-row-step, %.0f per SMSP-step\n", warps_per_smsp, (double)h[0] / 200, (double)h[0] / 200 / warps_per_smsp);
    exp_phase_kernel<true><<<1, 128 * warps_per_smsp>>>(out, cyc, 200); get();
    printf("exp phase (128 elements/thread, 25%% poly)   warps/SMSP %d: %.0f cycles per warp-row-step, %.0f per SMSP-step\n", warps_per_smsp, (double)h[0] / 200, (double)h[0] / 200 / warps_per_smsp);
  }
  for (int warps : {1, 4, 8}) {
    for (int st : {0, 1}) {
      tmem_ldst_kernel<<<1, 32 * warps>>>(out, cyc, 500, st); get();
      printf("TMEM %s 4 x (32x32b.x32) + wait, %d warps: %.0f cycles per round (16 KiB per warp) = %.1f B/clk/SM\n", st ? "st" : "ld", warps, (double)h[0] / 500, 16384.0 * warps / ((double)h[0] / 500));
    }
  }
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const char* mk[] = {"SS 128x128x64 (4 MMA)", "SS 128x64x64 (4 MMA)", "TS 128x64x128 (8 MMA)", "TS 128x64x64 (4 MMA)"};
  for (int kind = 0; kind < 4; ++kind)
    for (int ng : {1, 8})
      for (int cl : {0, 1}) {
        mma_kernel<<<1, 256, 80 * 1024>>>(cyc, kind, ng, cl, out); get();
        printf("MMA %-22s x%d%s: issue %lld cycles, issue+complete %lld cycles (%.0f per group)%s\n", mk[kind], ng, cl ? " + 4 warps of tcgen05.ld" : "", h[0], h[1],
               (double)h[1] / ng, cl ? "" : "");
        if (cl) printf("      concurrent tcgen05.ld rounds per warp: %lld (%.0f cycles each)\n", h[2], h[2] ? (double)h[1] / h[2] : 0.0);
      }
  cudaFuncSetAttribute(mma_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int ts : {0, 1})
    for (int bmn : {0, 1})
      for (int N : {32, 64, 128, 256})
        for (int nacc : {1, 2}) {
          if (nacc * N > 384) continue;
          mma_sweep_kernel<<<1, 64, 100 * 1024>>>(cyc, N, ts, bmn, nacc); get();
          printf("MMA sweep M=128 N=%3d K=16 %s B %s-major, %d accumulator(s): %.1f cycles per instruction (%.0f MAC/clk)\n", N, ts ? "A=TMEM" : "A=smem",
                 bmn ? "MN" : "K", nacc, (double)h[0] / 32, 128.0 * N * 16 / ((double)h[0] / 32));
        }
  cudaFuncSetAttribute(mma_sweep_uniform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int mode : {0, 1})
    for (int ts : {0, 1})
      for (int N : {32, 64, 128, 256}) {
        mma_sweep_uniform_kernel<<<1, 64, 100 * 1024>>>(cyc, N, ts, mode); get();
        printf("MMA uniform-issue (%s) M=128 N=%3d K=16 %s: %.1f cycles per instruction incl. completion, %.1f issue only (%.0f MAC/clk)\n",
               mode ? "warp loops, elect per group" : "elect once", N, ts ? "A=TMEM" : "A=smem", (double)h[0] / 32, (double)h[1] / 32, 128.0 * N * 16 / ((double)h[0] / 32));
      }
  pingpong_kernel<<<1, 64>>>(cyc, 1000); get();
  printf("mbarrier ping-pong between two warps: %.0f cycles per round trip\n", (double)h[0] / 1000);
  return 0;
}
