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

template <int KIND, typename TO, int NV>
__device__ __forceinline__ void epi_apply(const EpiArgs& e, long long row, int col, float (&v)[NV]) {
  if constexpr (KIND == EPI_STORE) {
    if (e.bias) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += __ldg(e.bias + col + i);
    }
    store_vec<NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + col, v);
  } else if constexpr (KIND == EPI_ROTARY) {
    const int pos = (int)(row % e.seq_len);
    const int half = e.dim_head >> 1;
    const float* sp = e.rot_sin + (long long)pos * half;
    const float* cp = e.rot_cos + (long long)pos * half;
    float o[NV];
#pragma unroll
    for (int i = 0; i < NV; i += 2) {
      const int j = ((col + i) % e.dim_head) >> 1;
      const float s = __ldg(sp + j), c = __ldg(cp + j);
      o[i] = v[i] * c - v[i + 1] * s;
      o[i + 1] = v[i + 1] * c + v[i] * s;
    }
    store_vec<NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + col, o);
  } else if constexpr (KIND == EPI_RESIDUAL) {
    // out = residual_in + acc + bias; residual_in = aux (fp32, ldaux) when given, else out itself (in place)
    float* p = reinterpret_cast<float*>(e.out) + row * e.ldo + col;
    const float* pin = e.aux ? reinterpret_cast<const float*>(e.aux) + row * e.ldaux + col : p;
    float r[NV];
    load_vec<NV>(pin, r);
#pragma unroll
    for (int i = 0; i < NV; ++i) r[i] += v[i] + (e.bias ? __ldg(e.bias + col + i) : 0.f);
    store_vec<NV>(p, r);
  } else if constexpr (KIND == EPI_GLU) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += __ldg(e.bias + col + i);
    store_vec<NV>(reinterpret_cast<TO*>(e.out2) + row * e.ldo2 + col, v);
    TO* po = reinterpret_cast<TO*>(e.out) + row * e.ldo + (col >> 1);
    if constexpr (NV >= 16) {
      float o[NV / 2];
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) o[i] = v[2 * i] * gelu_tanh(v[2 * i + 1]);
      store_vec<NV / 2>(po, o);
    } else {
#pragma unroll
      for (int i = 0; i < NV / 2; ++i) po[i] = from_f32<TO>(v[2 * i] * gelu_tanh(v[2 * i + 1]));
    }
  } else if constexpr (KIND == EPI_GELU) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] += __ldg(e.bias + col + i);
    store_vec<NV>(reinterpret_cast<TO*>(e.out2) + row * e.ldo2 + col, v);
    float o[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) o[i] = gelu_tanh(v[i]);
    store_vec<NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + col, o);
  } else if constexpr (KIND == EPI_GLU_BWD) {
    // acc column c is d(h[c]); pre-activations of (value, gate) sit at aux[2c], aux[2c+1]
    const TO* pa = reinterpret_cast<const TO*>(e.aux) + row * e.ldaux + 2 * col;
    TO* po = reinterpret_cast<TO*>(e.out) + row * e.ldo + 2 * col;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float u[NV], o[NV];
      load_vec<NV>(pa + h * NV, u);
#pragma unroll
      for (int i = 0; i < NV; i += 2) {
        const float dh = v[h * (NV / 2) + (i >> 1)];
        o[i] = dh * gelu_tanh(u[i + 1]);
        o[i + 1] = dh * u[i] * gelu_tanh_grad(u[i + 1]);
      }
      store_vec<NV>(po + h * NV, o);
    }
  } else if constexpr (KIND == EPI_GELU_BWD) {
    float u[NV];
    load_vec<NV>(reinterpret_cast<const TO*>(e.aux) + row * e.ldaux + col, u);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= gelu_tanh_grad(u[i]);
    store_vec<NV>(reinterpret_cast<TO*>(e.out) + row * e.ldo + col, v);
  } else if constexpr (KIND == EPI_ACCUM) {
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
