// ipx_contract.cuh -- the one definition of the separable cubic contraction, shared by every
// evaluation kernel (generic and tiled), so a query's result does not depend on the kernel,
// the table layout or its neighbours in the query stream.
//
// Reference: the dense form  coef = A @ F;  fq = sum coef[a,b,c] ttx[a] tty[b] ttz[c]
// (interpax/_spline.py:587-589, 796-800, 1094-1099) with A = A_CUBIC (x) A_CUBIC (x) A_CUBIC
// (interpax/_coefs.py:8-163; identity checked in tests/test_oracle_pins.py).  Here: per-axis
// weights w = tt . A_CUBIC (hermite_weights), then one node LAYER (position along the last
// axis) at a time: contract the other axes into (value part, last-axis-slope part), then
// combine the two layers with the last-axis weights.
#pragma once

#include "ipx_device.cuh"

namespace ipx {

// 16-byte shared-memory accesses
template <typename T, int NT> __device__ __forceinline__ void lds_record(const char* p, T* r) {
#pragma unroll
  for (int c = 0; c < (NT * (int)sizeof(T)) / 16; ++c) {
    float4 v = *reinterpret_cast<const float4*>(p + 16 * c);
    const T* s = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int k = 0; k < 16 / (int)sizeof(T); ++k) r[c * (16 / (int)sizeof(T)) + k] = s[k];
  }
}

template <typename T, int NT> __device__ __forceinline__ void ldg_record(const T* p, T* r) {
  if (NT == 2) ldg_rec2(p, r);
  if (NT == 4) ldg_rec4(p, r);
  if (NT == 8) ldg_rec8(p, r);
}

// 256-bit read-only global loads (sm_100: LDG.E.256): half the load instructions and L1 requests
// of the direct gathers.  The record must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const double* p, double* r) {
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
}
__device__ __forceinline__ void ldg256(const float* p, float* r) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
      : "l"(p));
}
template <typename T, int NT> __device__ __forceinline__ void ldg_record_wide(const T* p, T* r) {
  constexpr int BYTES = NT * (int)sizeof(T);
  if constexpr (BYTES == 64) {
    ldg256(p, r);
    ldg256(p + NT / 2, r + NT / 2);
  } else if constexpr (BYTES == 32) {
    ldg256(p, r);
  } else {
    ldg_record<T, NT>(p, r);
  }
}

// weights of NQ queries that walk a layer together
template <typename T, int ND, int NQ> struct WeightRefs {
  const T (*w[NQ])[4];
};

// One layer of the separable contraction, for NQ queries that share the layer's records.
// load(c, r) fills the record of column c (bit ax of c = upper corner along column axis ax) in
// canonical order (last axis fastest).  On return A[k][0] is the layer's value part and
// A[k][1] its last-axis-slope part for query k.
template <typename T, int ND, int NQ, typename Loader>
__device__ __forceinline__ void layer_contract(Loader load, const WeightRefs<T, ND, NQ>& W, T (&A)[NQ][2]) {
  if (ND == 1) {
    T r[2];
    load(0, r);
#pragma unroll
    for (int k = 0; k < NQ; ++k) { A[k][0] = r[0]; A[k][1] = r[1]; }
  } else if (ND == 2) {
#pragma unroll
    for (int k = 0; k < NQ; ++k) { A[k][0] = T(0); A[k][1] = T(0); }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      T r[4];  // [f, fy, fx, fxy]: index = (d/dy) + 2 (d/dx)
      load(ii, r);
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        A[k][0] = fma(r[2], W.w[k][0][2 + ii], fma(r[0], W.w[k][0][ii], A[k][0]));
        A[k][1] = fma(r[3], W.w[k][0][2 + ii], fma(r[1], W.w[k][0][ii], A[k][1]));
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < NQ; ++k) { A[k][0] = T(0); A[k][1] = T(0); }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      T u[NQ][4];  // index = (d/dz) + 2 (d/dx), y contracted
#pragma unroll
      for (int k = 0; k < NQ; ++k) { u[k][0] = T(0); u[k][1] = T(0); u[k][2] = T(0); u[k][3] = T(0); }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        T r[8];  // [f, fz, fy, fyz, fx, fxz, fxy, fxyz]: index = (d/dz) + 2 (d/dy) + 4 (d/dx)
        load(ii | (jj << 1), r);
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int src = (m & 1) + 4 * (m >> 1);
            u[k][m] = fma(r[src + 2], W.w[k][1][2 + jj], fma(r[src], W.w[k][1][jj], u[k][m]));
          }
        }
      }
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        A[k][0] = fma(u[k][2], W.w[k][0][2 + ii], fma(u[k][0], W.w[k][0][ii], A[k][0]));
        A[k][1] = fma(u[k][3], W.w[k][0][2 + ii], fma(u[k][1], W.w[k][0][ii], A[k][1]));
      }
    }
  }
}

}  // namespace ipx
