// GEMM epilogues shared by the tcgen05 GEMM (gemm_tc.cu) and the fp32 SIMT GEMM (gemm_simt.cu).
// A thread hands over NV consecutive accumulator columns of one output row; the epilogue fuses what the
// reference does right after the matmul (bias, rotary, residual add, GELU / GLU, their backward forms).
#pragma once
#include "common.cuh"

enum EpiKind : int {
  EPI_STORE = 0,      // out = acc (+ bias[col])
  EPI_ROTARY = 1,     // out = rotary(acc)            qkv projection, progen.py:83-87 (rotary on q, k AND v)
  EPI_RESIDUAL = 2,   // out(f32) += acc + bias       to_out / proj_out + residual, progen.py:103,148,230-231
  EPI_GLU = 3,        // out2 = pre-activation (interleaved value,gate), out = value * gelu(gate)   progen.py:139-141
  EPI_GELU = 4,       // out2 = pre-activation, out = gelu(pre)                                     progen.py:143
  EPI_GLU_BWD = 5,    // acc = d(out of GLU); aux = saved pre-activation; out = d(pre) interleaved
  EPI_GELU_BWD = 6,   // acc = d(gelu out); aux = saved pre-activation; out = acc * gelu'(pre)
  EPI_ACCUM = 7,      // out(f32) += acc  (atomic when several CTAs own the same tile; optional tril mask)
  EPI_NUM_KINDS = 8
};

struct EpiArgs {
  void* out; long long ldo;
  void* out2; long long ldo2;
  const float* bias;                 // [N] (already interleaved for GLU) or nullptr
  const void* aux; long long ldaux;  // saved pre-activations for the *_BWD kinds
  const float* rot_sin;              // [seq_len, dim_head/2]
  const float* rot_cos;
  int seq_len; int dim_head;
  int atomic;                        // EPI_ACCUM: use red.global.add
  int tril;                          // EPI_ACCUM: keep only col <= (row % tril_rows)
  int tril_rows;
};

// ---------------------------------------------------------------------------------------------------------
// How an epilogue thread moves its row slice to / from global memory.
//   DirectIO      : per-thread vector accesses (CUDA-core GEMM: 8 consecutive columns per thread)
//   WarpStagedIO  : the tcgen05 epilogue owns one ROW per lane (TMEM lane == row), so direct stores would touch 32
//                   different 128-byte lines per instruction (LSU wavefront bound).  Instead the warp transposes 64-byte
//                   row slices through a private shared-memory buffer so each instruction moves 8 rows x 64 contiguous
//                   bytes; same for the loads of residual / saved pre-activations.
struct DirectIO {
  template <int N, typename T> __device__ __forceinline__ void store(T* p, long long, const float (&v)[N], bool valid) const {
    if (valid) store_vec<N>(p, v);
  }
  template <int N, typename T> __device__ __forceinline__ void load(const T* p, long long, float (&v)[N], bool valid) const {
    if (valid) load_vec<N>(p, v);
    else {
#pragma unroll
      for (int i = 0; i < N; ++i) v[i] = 0.f;
    }
  }
};

constexpr int STAGE_WARP_BYTES = 32 * 64;                // 32 rows x (at most) 64 B per epilogue warp, 128 B aligned

// Byte offset of 16-byte piece q of row r in a warp's staging buffer whose rows hold PPR pieces (dense, no padding).
// The piece index is XOR-swizzled with the 128-byte line the row sits in, so that BOTH access patterns are
// bank-conflict free: lane = row (8 consecutive rows, one piece) and lane = (row, piece) in row-major order.
template <int PPR> __device__ __forceinline__ int stage_off(int r, int q) {
  return r * (PPR * 16) + ((q ^ ((r * PPR / 8) & (PPR - 1))) << 4);
}

struct WarpStagedIO {
  uint8_t* buf;          // this warp's staging buffer (generic pointer into shared memory)
  int lane;
  unsigned valid_mask;   // bit r: row r of this warp's 32-row block is inside the matrix

  template <typename T> static __device__ __forceinline__ uint4 pack16(const float* v) {
    uint4 t;
    if constexpr (sizeof(T) == 2) {
      t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]); t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
    } else {
      t.x = __float_as_uint(v[0]); t.y = __float_as_uint(v[1]); t.z = __float_as_uint(v[2]); t.w = __float_as_uint(v[3]);
    }
    return t;
  }
  // by value: a reference to shared memory makes the compiler read the four words with separate 4-byte LDS
  template <typename T> static __device__ __forceinline__ void unpack16(const uint4 t, float* v) {
    if constexpr (sizeof(T) == 2) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
    } else {
      v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
    }
  }

  // p: this lane's (row, col) element; rows of the warp are consecutive, `ld` elements apart
  template <int N, typename T> __device__ __forceinline__ void store(T* p, long long ld, const float (&v)[N], bool) const {
    constexpr int BYTES = N * (int)sizeof(T);
    constexpr int SW = BYTES < 64 ? BYTES : 64;          // slice width in bytes per pass
    constexpr int PPR = SW / 16;                         // 16-byte pieces per row slice
    constexpr int EPP = 16 / (int)sizeof(T);             // elements per piece
    uint8_t* base = reinterpret_cast<uint8_t*>(p - (long long)lane * ld);
#pragma unroll
    for (int s = 0; s < BYTES / SW; ++s) {
#pragma unroll
      for (int q = 0; q < PPR; ++q)
        *reinterpret_cast<uint4*>(buf + stage_off<PPR>(lane, q)) = pack16<T>(&v[s * (SW / (int)sizeof(T)) + q * EPP]);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < PPR; ++it) {
        const int idx = it * 32 + lane;
        const int r = idx / PPR, q = idx % PPR;
        if ((valid_mask >> r) & 1u) {
          const uint4 t = *reinterpret_cast<const uint4*>(buf + stage_off<PPR>(r, q));
          *reinterpret_cast<uint4*>(base + (long long)r * ld * (int)sizeof(T) + s * SW + q * 16) = t;
        }
      }
      __syncwarp();
    }
  }
  // load = load_issue (coalesced global loads into raw registers; `p` is only used for address arithmetic, `mask` says
  // which of the warp's 32 rows exist) + load_finish (transpose through the staging buffer).  (Issuing the next chunk's
  // loads between the two halves was measured and is slower: the bytes in flight stay capped by registers.  The CTA-pair
  // kernel gemm_tc2.cu stages this operand with TMA instead.)
  template <int N, typename T>
  __device__ __forceinline__ void load_issue(const T* p, long long ld, uint4 (&t)[N * (int)sizeof(T) / 16], unsigned mask) const {
    constexpr int BYTES = N * (int)sizeof(T);
    constexpr int SW = BYTES < 64 ? BYTES : 64;
    constexpr int PPR = SW / 16;
    constexpr int PASSES = BYTES / SW;
    const uint8_t* base = reinterpret_cast<const uint8_t*>(p - (long long)lane * ld);
#pragma unroll
    for (int s = 0; s < PASSES; ++s) {
#pragma unroll
      for (int it = 0; it < PPR; ++it) {
        const int idx = it * 32 + lane;
        const int r = idx / PPR, q = idx % PPR;
        t[s * PPR + it] = make_uint4(0u, 0u, 0u, 0u);
        if ((mask >> r) & 1u)
          t[s * PPR + it] = *reinterpret_cast<const uint4*>(base + (long long)r * ld * (int)sizeof(T) + s * SW + q * 16);
      }
    }
  }
  template <int N, typename T> __device__ __forceinline__ void load(const T* p, long long ld, float (&v)[N], bool) const {
    uint4 t[N * (int)sizeof(T) / 16];
    load_issue<N, T>(p, ld, t, valid_mask);
    load_finish<N, T>(t, v);
  }
  template <int N, typename T>
  __device__ __forceinline__ void load_finish(const uint4 (&t)[N * (int)sizeof(T) / 16], float (&v)[N]) const {
    constexpr int BYTES = N * (int)sizeof(T);
    constexpr int SW = BYTES < 64 ? BYTES : 64;
    constexpr int PPR = SW / 16;
    constexpr int EPP = 16 / (int)sizeof(T);
    constexpr int PASSES = BYTES / SW;
#pragma unroll
    for (int s = 0; s < PASSES; ++s) {
#pragma unroll
      for (int it = 0; it < PPR; ++it) {
        const int idx = it * 32 + lane;
        const int r = idx / PPR, q = idx % PPR;
        *reinterpret_cast<uint4*>(buf + stage_off<PPR>(r, q)) = t[s * PPR + it];
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < PPR; ++q)
        unpack16<T>(*reinterpret_cast<const uint4*>(buf + stage_off<PPR>(lane, q)), &v[s * (SW / (int)sizeof(T)) + q * EPP]);
      __syncwarp();
    }
  }
};

// v[i] += bias[col + i]: the same NV floats for every lane (broadcast), fetched with 16-byte loads
template <int NV> __device__ __forceinline__ void add_bias(const float* __restrict__ bias, int col, float (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; i += 4) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + i));
    v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
  }
}

// Every lane of the calling warp must enter (staged IO is warp-cooperative); `valid` says whether this lane's row exists.
template <int KIND, typename TO, int NV, typename IO>
__device__ __forceinline__ void epi_apply(const EpiArgs& e, const IO& io, long long row, int col, float (&v)[NV], bool valid) {
  constexpr bool FAST = sizeof(TO) == 2;      // bf16 outputs: hardware tanh is below the output rounding
  if constexpr (KIND == EPI_STORE) {
    if (e.bias) add_bias<NV>(e.bias, col, v);
    io.template store<NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + col, e.ldo, v, valid);
  } else if constexpr (KIND == EPI_ROTARY) {
    const int pos = (int)(row % e.seq_len);
    const int half = e.dim_head >> 1;
    const float* sp = e.rot_sin + (long long)pos * half;
    const float* cp = e.rot_cos + (long long)pos * half;
    float o[NV];
    if (e.dim_head % NV == 0) {
      // the NV columns sit inside one head: NV/2 consecutive (sin, cos) entries, 16-byte vector loads
      const int j0 = (col % e.dim_head) >> 1;
      float s[NV / 2], c[NV / 2];
      if constexpr (NV >= 32) {
        // consecutive rows are consecutive positions (a warp's 32 rows never straddle a sequence: seq_len % 32 == 0
        // is checked by the launcher), so the table slices form a [32 x NV/2] block: coalesced staged loads
        io.template load<NV / 2>(sp + j0, half, s, true);
        io.template load<NV / 2>(cp + j0, half, c, true);
      } else {
        load_vec<NV / 2>(sp + j0, s);
        load_vec<NV / 2>(cp + j0, c);
      }
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        o[i] = v[i] * c[i >> 1] - v[i + 1] * s[i >> 1];
        o[i + 1] = v[i + 1] * c[i >> 1] + v[i] * s[i >> 1];
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        const int j = ((col + i) % e.dim_head) >> 1;
        const float s = __ldg(sp + j), c = __ldg(cp + j);
        o[i] = v[i] * c - v[i + 1] * s;
        o[i + 1] = v[i + 1] * c + v[i] * s;
      }
    }
    io.template store<NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + col, e.ldo, o, valid);
  } else if constexpr (KIND == EPI_RESIDUAL) {
    // out = residual_in + acc + bias; residual_in = aux (fp32, ldaux) when given, else out itself (in place)
    float* p = reinterpret_cast<float*>(e.out) + row * e.ldo + col;
    float r[NV];
    if (e.aux) io.template load<NV>(reinterpret_cast<const float*>(e.aux) + row * e.ldaux + col, e.ldaux, r, valid);
    else io.template load<NV>(const_cast<const float*>(p), e.ldo, r, valid);
    if (e.bias) add_bias<NV>(e.bias, col, v);
#pragma unroll
    for (int i = 0; i < NV; ++i) r[i] += v[i];
    io.template store<NV>(p, e.ldo, r, valid);
  } else if constexpr (KIND == EPI_GLU) {
    add_bias<NV>(e.bias, col, v);
    io.template store<NV>(reinterpret_cast<TO*>(e.out2) + row * e.ldo2 + col, e.ldo2, v, valid);
    TO* po = reinterpret_cast<TO*>(e.out) + row * e.ldo + (col >> 1);
    if constexpr (NV >= 16) {
      float o[NV / 2];
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) o[i] = v[2 * i] * gelu_fwd<FAST>(v[2 * i + 1]);
      io.template store<NV / 2>(po, e.ldo, o, valid);
    } else {
      if (valid) {
#pragma unroll
        for (int i = 0; i < NV / 2; ++i) po[i] = from_f32<TO>(v[2 * i] * gelu_fwd<FAST>(v[2 * i + 1]));
      }
    }
  } else if constexpr (KIND == EPI_GELU) {
    add_bias<NV>(e.bias, col, v);
    io.template store<NV>(reinterpret_cast<TO*>(e.out2) + row * e.ldo2 + col, e.ldo2, v, valid);
    float o[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) o[i] = gelu_fwd<FAST>(v[i]);
    io.template store<NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + col, e.ldo, o, valid);
  } else if constexpr (KIND == EPI_GLU_BWD) {
    // acc column c is d(h[c]); pre-activations of (value, gate) sit at aux[2c], aux[2c+1]; one 2*NV-wide load / store
    float u[2 * NV];
    io.template load<2 * NV>(reinterpret_cast<const TO*>(e.aux) + row * e.ldaux + 2 * col, e.ldaux, u, valid);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float dh = v[i], val = u[2 * i], gate = u[2 * i + 1];
      float gf, gd;
      gelu_fwd_bwd<FAST>(gate, gf, gd);
      u[2 * i] = dh * gf;
      u[2 * i + 1] = dh * val * gd;
    }
    io.template store<2 * NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + 2 * col, e.ldo, u, valid);
  } else if constexpr (KIND == EPI_GELU_BWD) {
    float u[NV];
    io.template load<NV>(reinterpret_cast<const TO*>(e.aux) + row * e.ldaux + col, e.ldaux, u, valid);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= gelu_bwd<FAST>(u[i]);
    io.template store<NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + col, e.ldo, v, valid);
  } else if constexpr (KIND == EPI_ACCUM) {
    if (!valid) return;
    float* p = reinterpret_cast<float*>(e.out) + row * e.ldo + col;
    const int lim = e.tril ? (int)(row % e.tril_rows) : 0x7fffffff;
    if (e.atomic) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        if (col + i + 3 <= lim) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "f"(v[i]), "f"(v[i + 1]),
                       "f"(v[i + 2]), "f"(v[i + 3]) : "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (col + i + j <= lim) atomicAdd(p + i + j, v[i + j]);
        }
      }
    } else {
      float r[NV];
      load_vec<NV>(p, r);
#pragma unroll
      for (int i = 0; i < NV; ++i) r[i] += (col + i <= lim) ? v[i] : 0.f;
      store_vec<NV>(p, r);
    }
  }
}

// kinds whose epilogue reads a second [rows x cols] operand from global memory (residual stream / saved pre-activations)
template <int KIND> constexpr bool epi_has_aux = KIND == EPI_RESIDUAL || KIND == EPI_GLU_BWD || KIND == EPI_GELU_BWD;
