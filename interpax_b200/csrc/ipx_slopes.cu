// ipx_slopes.cu -- knot-slope precompute: approx_df as CUDA kernels.
//
// Replaces interpax/_fd_derivs.py:9-443 of the reference.  A C-contiguous table is
// viewed as (outer, n, inner) with the differentiated axis in the middle, so no
// moveaxis copy is made.  Local stencils (cubic, cardinal, monotonic, akima) are one
// thread per output element, coalesced along the fastest table dimension; cubic2 (a
// tridiagonal solve along the axis) is one thread per column with the knot-only part of
// the Thomas recurrence factored out and shared by all columns.
#include <cuda_runtime.h>

#include "ipx_device.cuh"
#include "ipx_internal.cuh"

namespace ipx {

constexpr int kSlopeThreads = 256;

template <typename T> struct SlopeArgs {
  const T* __restrict__ x;
  const T* __restrict__ f;
  T* __restrict__ out;
  int64_t outer, inner;
  int32_t n;
};

template <typename T>
__device__ __forceinline__ T secant(const SlopeArgs<T>& a, int64_t base, int k) {
  // s_k = (f[k+1] - f[k]) * where(dx == 0, 0, 1/dx)
  // products and sums through mul_rn / add_rn so that no caller contracts them into an FMA:
  // the per-table kernels and the fused one-pass kernel below must produce the same bits
  T df = a.f[base + (int64_t)(k + 1) * a.inner] - a.f[base + (int64_t)k * a.inner];
  return mul_rn(safe_recip(a.x[k + 1] - a.x[k]), df);
}

template <typename T> __device__ __forceinline__ T sign_of(T v) {
  return v > T(0) ? T(1) : (v < T(0) ? T(-1) : (v == T(0) ? T(0) : v));  // NaN stays NaN
}

template <typename T> __device__ __forceinline__ T safediv(T a, T b) {
  // _fd_derivs.py:427-443 with fill = 0, threshold = 0
  bool m = t_abs(b) <= T(0);
  return (m ? T(0) : a) / (m ? T(1) : b);
}

// decompose a flat element index of the (outer, n, inner) view
template <typename T>
__device__ __forceinline__ void split_index(const SlopeArgs<T>& a, int64_t e, int64_t& base, int& k) {
  int64_t i = e % a.inner;
  int64_t r = e / a.inner;
  k = (int)(r % a.n);
  int64_t o = r / a.n;
  base = o * a.n * a.inner + i;  // address of element k = 0 of this column
}

// ------------------------------------------------------------------ cubic  (_cubic1)
template <typename T> __global__ void slopes_cubic_kernel(const SlopeArgs<T> a) {
  const int64_t total = a.outer * a.n * a.inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t base; int k;
    split_index(a, e, base, k);
    T v;
    if (k == 0) v = secant(a, base, 0);
    else if (k == a.n - 1) v = secant(a, base, a.n - 2);
    else v = T(0.5) * add_rn(secant(a, base, k - 1), secant(a, base, k));
    a.out[e] = v;
  }
}

// ------------------------------------------------------------------ cardinal (_cardinal)
template <typename T> __global__ void slopes_cardinal_kernel(const SlopeArgs<T> a, T one_minus_c) {
  const int64_t total = a.outer * a.n * a.inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t base; int k;
    split_index(a, e, base, k);
    int lo = k == 0 ? 0 : k - 1;
    int hi = k == a.n - 1 ? a.n - 1 : k + 1;
    T df = a.f[base + (int64_t)hi * a.inner] - a.f[base + (int64_t)lo * a.inner];
    T v = (k == 0 || k == a.n - 1) ? df * safe_recip(a.x[hi] - a.x[lo])
                                   : safe_recip(a.x[hi] - a.x[lo]) * df;
    a.out[e] = one_minus_c * v;
  }
}

// ------------------------------------------------------------------ monotonic (_monotonic)
template <typename T>
__device__ __forceinline__ T pchip_edge(T h0, T h1, T m0, T m1) {
  // _fd_derivs.py:365-376
  T d = add_rn(mul_rn(T(2) * h0 + h1, m0), -mul_rn(h0, m1)) / (h0 + h1);
  bool mask = sign_of(d) != sign_of(m0);
  bool mask2 = (sign_of(m0) != sign_of(m1)) && (t_abs(d) > T(3) * t_abs(m0));
  bool mmm = (!mask) && mask2;
  if (mask) d = T(0);
  if (mmm) d = T(3) * m0;
  return d;
}

template <typename T> __global__ void slopes_monotonic_kernel(const SlopeArgs<T> a, int zero_ends) {
  const int64_t total = a.outer * a.n * a.inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int n = a.n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t base; int k;
    split_index(a, e, base, k);
    T v;
    if (k == 0 || k == n - 1) {
      if (zero_ends) {
        v = T(0);
      } else {
        // intervals 0,1 at the start; n-2,n-3 at the end; JAX clamps the second index
        // when there is only one interval (n == 2)
        int j0 = k == 0 ? 0 : n - 2;
        int j1 = k == 0 ? min(1, n - 2) : max(n - 3, 0);
        T hi0 = safe_recip(a.x[j0 + 1] - a.x[j0]);
        T hi1 = safe_recip(a.x[j1 + 1] - a.x[j1]);
        T m0 = hi0 * (a.f[base + (int64_t)(j0 + 1) * a.inner] - a.f[base + (int64_t)j0 * a.inner]);
        T m1 = hi1 * (a.f[base + (int64_t)(j1 + 1) * a.inner] - a.f[base + (int64_t)j1 * a.inner]);
        v = pchip_edge(T(1) / hi0, T(1) / hi1, m0, m1);  // hk = 1 / hki, _fd_derivs.py:378
      }
    } else {
      T h0 = a.x[k] - a.x[k - 1];      // hk[k-1]
      T h1 = a.x[k + 1] - a.x[k];      // hk[k]
      T fm = a.f[base + (int64_t)(k - 1) * a.inner];
      T f0 = a.f[base + (int64_t)k * a.inner];
      T fp = a.f[base + (int64_t)(k + 1) * a.inner];
      T m0 = safe_recip(h0) * (f0 - fm);
      T m1 = safe_recip(h1) * (fp - f0);
      bool cond = (sign_of(m1) != sign_of(m0)) || (m1 == T(0)) || (m0 == T(0));
      T w1 = T(2) * h1 + h0;
      T w2 = h1 + T(2) * h0;
      T whmean = safediv(safediv(w1, m0) + safediv(w2, m1), w1 + w2);
      v = cond ? T(0) : safediv(T(1), whmean);
    }
    a.out[e] = v;
  }
}

// ------------------------------------------------------------------ akima (_akima)
// m[j], j in [0, n+2]: extended secant array of _fd_derivs.py:397-409 (zero-initialised,
// which matters only for n == 2).
template <typename T>
__device__ __forceinline__ T akima_secant(const SlopeArgs<T>& a, int64_t base, int k) {
  T dx = a.x[k + 1] - a.x[k];
  bool z = dx == T(0);
  T dxi = z ? T(0) : T(1) / dx;
  return (a.f[base + (int64_t)(k + 1) * a.inner] - a.f[base + (int64_t)k * a.inner]) * dxi;
}

template <typename T>
__device__ __forceinline__ T akima_m(const SlopeArgs<T>& a, int64_t base, int j) {
  const int n = a.n;
  if (j >= 2 && j <= n) return akima_secant(a, base, j - 2);
  if (n == 2) {
    // only one true secant s0 = m[2]; the reference reads m[3] before it is written
    T s0 = akima_secant(a, base, 0);
    T m1 = T(2) * s0 - T(0);
    if (j == 1) return m1;
    if (j == 0) return T(2) * m1 - s0;
    T m3 = T(2) * s0 - m1;        // m[-2] = 2 m[-3] - m[-4]
    if (j == 3) return m3;
    return T(2) * m3 - s0;        // m[-1] = 2 m[-2] - m[-3]
  }
  if (j < 2) {
    T m2 = akima_secant(a, base, 0), m3 = akima_secant(a, base, 1);
    T m1 = T(2) * m2 - m3;
    return j == 1 ? m1 : T(2) * m1 - m2;
  }
  T mn = akima_secant(a, base, n - 2), mn1 = akima_secant(a, base, n - 3);  // m[n], m[n-1]
  T mp1 = T(2) * mn - mn1;                                                  // m[n+1]
  return j == n + 1 ? mp1 : T(2) * mp1 - mn;
}

// max over non-negative values AND NaN: jnp.max (_fd_derivs.py:421) propagates a NaN, which makes
// the threshold NaN and sends every node to the fall-back slope.  A NaN is stored as the positive
// quiet NaN, whose bit pattern compares above every finite value and +inf.
__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  atomicMax(reinterpret_cast<int*>(addr), v != v ? 0x7fc00000 : __float_as_int(v));
}
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
  atomicMax(reinterpret_cast<unsigned long long*>(addr),
            v != v ? 0x7ff8000000000000ull : (unsigned long long)__double_as_longlong(v));
}

template <typename T> __global__ void akima_init_kernel(T* gmax) { *gmax = T(0); }

template <typename T>
__global__ void akima_max_kernel(const SlopeArgs<T> a, T* __restrict__ gmax) {
  const int64_t total = a.outer * a.n * a.inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  T local = T(0);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t base; int k;
    split_index(a, e, base, k);
    T m1 = akima_m(a, base, k), m2 = akima_m(a, base, k + 1);
    T m3 = akima_m(a, base, k + 2), m4 = akima_m(a, base, k + 3);
    T f12 = t_abs(m4 - m3) + t_abs(m2 - m1);
    if (f12 > local || f12 != f12) local = f12;  // (a NaN sticks: nothing compares above it)
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    T o = __shfl_xor_sync(0xffffffffu, local, off);
    if (o > local || o != o) local = o;
  }
  if ((threadIdx.x & 31) == 0 && (local > T(0) || local != local)) atomic_max_nonneg(gmax, local);
}

template <typename T>
__global__ void slopes_akima_kernel(const SlopeArgs<T> a, const T* __restrict__ gmax) {
  const int64_t total = a.outer * a.n * a.inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const T thresh = T(1e-9) * (*gmax);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t base; int k;
    split_index(a, e, base, k);
    T m1 = akima_m(a, base, k), m2 = akima_m(a, base, k + 1);
    T m3 = akima_m(a, base, k + 2), m4 = akima_m(a, base, k + 3);
    T m4m3 = t_abs(m4 - m3), m2m1 = t_abs(m2 - m1);
    T f12 = m4m3 + m2m1;
    bool mask = f12 > thresh;
    T df = (m4m3 * m2 + m2m1 * m3) / (mask ? f12 : T(1));
    a.out[e] = mask ? df : T(0.5) * (m4 + m1);  // the code's 0.5*(m[3:]+m[:-3]), _fd_derivs.py:423
  }
}

// Complex data (IPX_SLOPE_AKIMA_COMPLEX): `inner` counts REAL elements and is even, every
// (re, im) pair is one complex value.  The extended secants are linear, so each part goes through
// akima_m on its own; only the weights |m4 - m3|, |m2 - m1| (jnp.abs of a complex difference,
// _fd_derivs.py:414) and the global maximum of their sum couple the two parts.
template <typename T> __device__ __forceinline__ T t_hypot(T a, T b);
template <> __device__ __forceinline__ float t_hypot<float>(float a, float b) { return hypotf(a, b); }
template <> __device__ __forceinline__ double t_hypot<double>(double a, double b) { return hypot(a, b); }

template <typename T>
__device__ __forceinline__ void split_index_pairs(const SlopeArgs<T>& a, int64_t e, int64_t& base, int& k) {
  const int64_t half = a.inner >> 1;
  int64_t i = e % half;
  int64_t r = e / half;
  k = (int)(r % a.n);
  int64_t o = r / a.n;
  base = o * a.n * a.inner + 2 * i;  // real part of element k = 0 of this column
}

template <typename T> struct AkimaPair {
  T w43, w21, m1[2], m2[2], m3[2], m4[2];
};
template <typename T>
__device__ __forceinline__ void akima_pair(const SlopeArgs<T>& a, int64_t base, int k, AkimaPair<T>& p) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    p.m1[c] = akima_m(a, base + c, k);
    p.m2[c] = akima_m(a, base + c, k + 1);
    p.m3[c] = akima_m(a, base + c, k + 2);
    p.m4[c] = akima_m(a, base + c, k + 3);
  }
  p.w43 = t_hypot(p.m4[0] - p.m3[0], p.m4[1] - p.m3[1]);
  p.w21 = t_hypot(p.m2[0] - p.m1[0], p.m2[1] - p.m1[1]);
}

template <typename T>
__global__ void akima_max_complex_kernel(const SlopeArgs<T> a, T* __restrict__ gmax) {
  const int64_t total = a.outer * a.n * (a.inner >> 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  T local = T(0);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t base; int k;
    split_index_pairs(a, e, base, k);
    AkimaPair<T> p;
    akima_pair(a, base, k, p);
    T f12 = p.w43 + p.w21;
    if (f12 > local || f12 != f12) local = f12;  // (a NaN sticks: nothing compares above it)
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    T o = __shfl_xor_sync(0xffffffffu, local, off);
    if (o > local || o != o) local = o;
  }
  if ((threadIdx.x & 31) == 0 && (local > T(0) || local != local)) atomic_max_nonneg(gmax, local);
}

template <typename T>
__global__ void slopes_akima_complex_kernel(const SlopeArgs<T> a, const T* __restrict__ gmax) {
  const int64_t total = a.outer * a.n * (a.inner >> 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const T thresh = T(1e-9) * (*gmax);
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t base; int k;
    split_index_pairs(a, e, base, k);
    AkimaPair<T> p;
    akima_pair(a, base, k, p);
    const T f12 = p.w43 + p.w21;
    const bool mask = f12 > thresh;
    const T den = mask ? f12 : T(1);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      T df = (p.w43 * p.m2[c] + p.w21 * p.m3[c]) / den;
      a.out[base + (int64_t)k * a.inner + c] = mask ? df : T(0.5) * (p.m4[c] + p.m1[c]);
    }
  }
}

// ------------------------------------------------------------------ cubic2 (_cubic2)
template <typename T> struct BcArgs {
  int32_t start, end;                       // IPX_BC_*
  const T* __restrict__ start_value;        // per column or NULL (zeros)
  const T* __restrict__ end_value;
};

// row k of the tridiagonal system: sub-diagonal a_k (A[k,k-1]), diagonal b_k, super-diagonal
// c_k (A[k,k+1]) -- knots and boundary kinds only (_fd_derivs.py:208-213, 219-222, 231-239,
// 250-253, 262-270), then the duplicate-knot neutralisation (:281-285).
template <typename T>
__device__ __forceinline__ void cubic2_row(const T* __restrict__ x, int n, int bs, int be, int k,
                                           T& lower, T& diag, T& upper, bool& masked) {
  lower = T(0); upper = T(0);
  if (k == 0) {
    if (bs == IPX_BC_NOT_A_KNOT) { diag = x[2] - x[1]; upper = x[2] - x[0]; }
    else if (bs == IPX_BC_FIRST) { diag = T(1); upper = T(0); }
    else { T d0 = x[1] - x[0]; diag = T(2) * d0; upper = d0; }
  } else if (k == n - 1) {
    if (be == IPX_BC_NOT_A_KNOT) { diag = x[n - 2] - x[n - 3]; lower = x[n - 1] - x[n - 3]; }
    else if (be == IPX_BC_FIRST) { diag = T(1); lower = T(0); }
    else { T d1 = x[n - 1] - x[n - 2]; diag = T(2) * d1; lower = d1; }
  } else {
    T dl = x[k] - x[k - 1], dr = x[k + 1] - x[k];
    diag = T(2) * (dl + dr); upper = dl; lower = dr;
  }
  masked = diag == T(0);
  if (masked) { diag = T(1); lower = T(0); upper = T(0); }
}

// knot-only part of the Thomas recurrence (lineax Tridiagonal solver order):
//   den_k = b_k - a_k c'_{k-1};  c'_k = c_k / den_k
// one thread, sequential; ws[0..n) = c', ws[n..2n) = den
template <typename T>
__global__ void cubic2_factor_kernel(const T* __restrict__ x, int n, int bs, int be,
                                     T* __restrict__ ws) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  T cprev = T(0);
  for (int k = 0; k < n; ++k) {
    T lo, di, up;
    bool masked;
    cubic2_row(x, n, bs, be, k, lo, di, up, masked);
    T den = di - lo * cprev;
    cprev = up / den;
    ws[k] = cprev;
    ws[n + k] = den;
  }
}

template <typename T>
__device__ __forceinline__ T cubic2_rhs(const SlopeArgs<T>& a, const BcArgs<T>& bc, int64_t base,
                                        int64_t col, int k, bool masked) {
  const int n = a.n;
  const T* __restrict__ x = a.x;
  auto F = [&](int j) { return a.f[base + (int64_t)j * a.inner]; };
  auto dxr = [&](int j) { return x[j + 1] - x[j]; };
  auto s = [&](int j) {  // secant slope, with the reference's double `where` guard
    T d = dxr(j);
    T dxi = d == T(0) ? T(0) : T(1) / d;
    return dxi * (F(j + 1) - F(j));
  };
  if (masked) return T(0);  // _fd_derivs.py:286
  if (k == 0) {
    if (bc.start == IPX_BC_NOT_A_KNOT) {
      T d = x[2] - x[0];
      return ((dxr(0) + T(2) * d) * dxr(1) * s(0) + dxr(0) * dxr(0) * s(1)) / d;
    }
    T v = bc.start_value ? bc.start_value[col] : T(0);
    if (bc.start == IPX_BC_FIRST) return v;
    T d0 = dxr(0);
    return T(-0.5) * v * (d0 * d0) + T(3) * (F(1) - F(0));
  }
  if (k == n - 1) {
    if (bc.end == IPX_BC_NOT_A_KNOT) {
      T d = x[n - 1] - x[n - 3];
      return (dxr(n - 2) * dxr(n - 2) * s(n - 3) + (T(2) * d + dxr(n - 2)) * dxr(n - 3) * s(n - 2)) / d;
    }
    T v = bc.end_value ? bc.end_value[col] : T(0);
    if (bc.end == IPX_BC_FIRST) return v;
    T d1 = dxr(n - 2);
    return T(0.5) * v * (d1 * d1) + T(3) * (F(n - 1) - F(n - 2));
  }
  return T(3) * (dxr(k) * s(k - 1) + dxr(k - 1) * s(k));
}

// one thread per column: forward sweep d'_k = (d_k - a_k d'_{k-1}) / den_k into `out`,
// then back substitution x_k = d'_k - c'_k x_{k+1} in place.
template <typename T>
__global__ void cubic2_solve_kernel(const SlopeArgs<T> a, const BcArgs<T> bc,
                                    const T* __restrict__ ws) {
  const int64_t cols = a.outer * a.inner;
  const int n = a.n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < cols; col += stride) {
    int64_t o = col / a.inner, i = col - o * a.inner;
    int64_t base = o * n * a.inner + i;
    T dprev = T(0);
    for (int k = 0; k < n; ++k) {
      T lo, di, up;
      bool masked;
      cubic2_row(a.x, n, bc.start, bc.end, k, lo, di, up, masked);
      T rhs = cubic2_rhs(a, bc, base, col, k, masked);
      dprev = (rhs - lo * dprev) / ws[n + k];
      a.out[base + (int64_t)k * a.inner] = dprev;
    }
    T xnext = T(0);
    for (int k = n - 1; k >= 0; --k) {
      xnext = a.out[base + (int64_t)k * a.inner] - ws[k] * xnext;
      a.out[base + (int64_t)k * a.inner] = xnext;
    }
  }
}

// T y = rhs for a given right-hand side (forward mode of the slopes with respect to the knots:
// the tangent system has the same matrix), in place, one thread per column
template <typename T>
__global__ void cubic2_solve_rhs_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t outer, int64_t inner,
                                        int n, int bs, int be, const T* __restrict__ ws) {
  const int64_t cols = outer * inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < cols; col += stride) {
    const int64_t o = col / inner, i = col - o * inner;
    const int64_t base = o * n * inner + i;
    T dprev = T(0);
    for (int k = 0; k < n; ++k) {
      T lo, di, up;
      bool masked;
      cubic2_row(x, n, bs, be, k, lo, di, up, masked);
      const T rhs = masked ? T(0) : y[base + (int64_t)k * inner];
      dprev = (rhs - lo * dprev) / ws[n + k];
      y[base + (int64_t)k * inner] = dprev;
    }
    T xnext = T(0);
    for (int k = n - 1; k >= 0; --k) {
      xnext = y[base + (int64_t)k * inner] - ws[k] * xnext;
      y[base + (int64_t)k * inner] = xnext;
    }
  }
}

// ---- long axes: windowed Thomas ------------------------------------------------------
// The spline system is strictly diagonally dominant in its interior rows
// (|lower| + |upper| = diag / 2 for any knot spacing), so the influence of a row on the
// solution decays at least like (2 - sqrt(3))^distance = 0.268^distance.  A chunk of rows can
// therefore be solved from a window that extends kHalo = 64 rows beyond it (0.268^64 = 2e-37:
// far below one ulp) instead of from the whole axis: forward elimination enters the chunk
// through the left halo, a backward elimination through the right halo supplies the value
// just past the chunk, and back substitution covers the chunk.  Chunks at the ends of the
// axis include the true boundary rows.  This turns the 2 n sequential steps per column of the
// plain Thomas sweep into n / kChunk independent pieces per column (1e6 knots x 64 columns:
// 1.4 s -> milliseconds) with results equal to the sequential sweep up to rounding.
constexpr int kChunk = 256;
constexpr int kHalo = 64;

// matrix-only part for every chunk: c'_k of the chunk's own rows -> ws[k]
template <typename T>
__global__ void cubic2_wfactor_kernel(const T* __restrict__ x, int n, int bs, int be,
                                      T* __restrict__ ws) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = p * kChunk;
  if (s >= n) return;
  const int e = min(n, s + kChunk);
  T cprev = T(0);
  for (int k = max(0, s - kHalo); k < e; ++k) {
    T lo, di, up;
    bool masked;
    cubic2_row(x, n, bs, be, k, lo, di, up, masked);
    const T den = di - lo * cprev;
    cprev = up / den;
    if (k >= s) ws[k] = cprev;
  }
}

template <typename T>
__global__ void cubic2_wsolve_kernel(const SlopeArgs<T> a, const BcArgs<T> bc,
                                     const T* __restrict__ ws) {
  const int64_t cols = a.outer * a.inner;
  const int n = a.n;
  const int chunks = (n + kChunk - 1) / kChunk;
  const int64_t total = cols * chunks;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const int64_t col = t % cols;
    const int p = (int)(t / cols);
    const int64_t o = col / a.inner, i = col - o * a.inner;
    const int64_t base = o * n * a.inner + i;
    const int s = p * kChunk, e = min(n, s + kChunk);
    // forward elimination: left halo (nothing stored), then the chunk (d' -> out)
    T cprev = T(0), dprev = T(0);
    for (int k = max(0, s - kHalo); k < e; ++k) {
      T lo, di, up;
      bool masked;
      cubic2_row(a.x, n, bc.start, bc.end, k, lo, di, up, masked);
      const T rhs = cubic2_rhs(a, bc, base, col, k, masked);
      const T den = di - lo * cprev;
      cprev = up / den;
      dprev = (rhs - lo * dprev) / den;
      if (k >= s) a.out[base + (int64_t)k * a.inner] = dprev;
    }
    // value just past the chunk from a backward elimination through the right halo:
    //   x_e = dq - aq x_{e-1}   and   x_{e-1} = d'_{e-1} - c'_{e-1} x_e
    T xnext = T(0);
    if (e < n) {
      T aq = T(0), dq = T(0);
      for (int k = min(n, e + kHalo) - 1; k >= e; --k) {
        T lo, di, up;
        bool masked;
        cubic2_row(a.x, n, bc.start, bc.end, k, lo, di, up, masked);
        const T rhs = cubic2_rhs(a, bc, base, col, k, masked);
        const T den = di - up * aq;
        aq = lo / den;
        dq = (rhs - up * dq) / den;
      }
      xnext = (dq - aq * dprev) / (T(1) - aq * cprev);
    }
    for (int k = e - 1; k >= s; --k) {
      xnext = a.out[base + (int64_t)k * a.inner] - ws[k] * xnext;
      a.out[base + (int64_t)k * a.inner] = xnext;
    }
  }
}

// n == 2: both ends take the secant when not-a-knot (_fd_derivs.py:175-179), then the
// 2x2 system is solved by the same recurrence.  n == 3 with two not-a-knots: the parabola
// through the three points, a dense 3x3 solve with partial pivoting (:185-203).
template <typename T>
__global__ void cubic2_small_kernel(const SlopeArgs<T> a, const BcArgs<T> bc) {
  const int64_t cols = a.outer * a.inner;
  const int n = a.n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const T* __restrict__ x = a.x;
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < cols; col += stride) {
    int64_t o = col / a.inner, i = col - o * a.inner;
    int64_t base = o * n * a.inner + i;
    auto F = [&](int j) { return a.f[base + (int64_t)j * a.inner]; };
    if (n == 2) {
      T dx0 = x[1] - x[0];
      T s0 = (dx0 == T(0) ? T(0) : T(1) / dx0) * (F(1) - F(0));
      // rows: start row (diag, upper, rhs), end row (lower, diag, rhs)
      T d0, u0, r0, l1, d1, r1;
      if (bc.start == IPX_BC_NOT_A_KNOT) { d0 = T(1); u0 = T(0); r0 = s0; }
      else if (bc.start == IPX_BC_FIRST) { d0 = T(1); u0 = T(0); r0 = bc.start_value ? bc.start_value[col] : T(0); }
      else { d0 = T(2) * dx0; u0 = dx0; r0 = T(-0.5) * (bc.start_value ? bc.start_value[col] : T(0)) * (dx0 * dx0) + T(3) * (F(1) - F(0)); }
      if (bc.end == IPX_BC_NOT_A_KNOT) { d1 = T(1); l1 = T(0); r1 = s0; }
      else if (bc.end == IPX_BC_FIRST) { d1 = T(1); l1 = T(0); r1 = bc.end_value ? bc.end_value[col] : T(0); }
      else { d1 = T(2) * dx0; l1 = dx0; r1 = T(0.5) * (bc.end_value ? bc.end_value[col] : T(0)) * (dx0 * dx0) + T(3) * (F(1) - F(0)); }
      if (d0 == T(0)) { d0 = T(1); u0 = T(0); r0 = T(0); }
      if (d1 == T(0)) { d1 = T(1); l1 = T(0); r1 = T(0); }
      T den0 = d0;
      T dp0 = r0 / den0, cp0 = u0 / den0;
      T den1 = d1 - l1 * cp0;
      T dp1 = (r1 - l1 * dp0) / den1;
      a.out[base + a.inner] = dp1;
      a.out[base] = dp0 - cp0 * dp1;
    } else {  // n == 3, not-a-knot at both ends
      T dx0 = x[1] - x[0], dx1 = x[2] - x[1];
      T s0 = (dx0 == T(0) ? T(0) : T(1) / dx0) * (F(1) - F(0));
      T s1 = (dx1 == T(0) ? T(0) : T(1) / dx1) * (F(2) - F(1));
      T A[3][4] = {{T(1), T(1), T(0), T(2) * s0},
                   {dx1, T(2) * (dx0 + dx1), dx0, T(3) * (dx0 * s1 + dx1 * s0)},
                   {T(0), T(1), T(1), T(2) * s1}};
      // LU with partial pivoting (what jnp.linalg.solve does)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int piv = c;
#pragma unroll
        for (int r = c + 1; r < 3; ++r)
          if (t_abs(A[r][c]) > t_abs(A[piv][c])) piv = r;
        if (piv != c) {
#pragma unroll
          for (int j = 0; j < 4; ++j) { T t = A[c][j]; A[c][j] = A[piv][j]; A[piv][j] = t; }
        }
#pragma unroll
        for (int r = c + 1; r < 3; ++r) {
          T m = A[r][c] / A[c][c];
#pragma unroll
          for (int j = c; j < 4; ++j) A[r][j] -= m * A[c][j];
        }
      }
      T v2 = A[2][3] / A[2][2];
      T v1 = (A[1][3] - A[1][2] * v2) / A[1][1];
      T v0 = (A[0][3] - A[0][1] * v1 - A[0][2] * v2) / A[0][0];
      a.out[base] = v0;
      a.out[base + a.inner] = v1;
      a.out[base + 2 * a.inner] = v2;
    }
  }
}

template <typename T> __global__ void fill_zero_kernel(T* __restrict__ out, int64_t total) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride)
    out[e] = T(0);
}

static int slope_grid(int64_t total) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (total + kSlopeThreads - 1) / kSlopeThreads;
  int64_t cap = (int64_t)sms * 16;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

template <typename T>
static int launch_slopes(const T* x, int32_t n, const T* f, int64_t outer, int64_t inner,
                         int32_t method, const ipx_slope_opts* opts, T* out, void* workspace,
                         void* stream) {
  if (outer < 0 || inner < 0 || n < 0) return IPX_EINVAL;
  const int64_t total = outer * n * inner;
  if (total == 0) return IPX_OK;
  if (x == nullptr || f == nullptr || out == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SlopeArgs<T> a{x, f, out, outer, inner, n};
  const int grid = slope_grid(total);
  switch (method) {
    case IPX_SLOPE_ZERO:
      fill_zero_kernel<T><<<grid, kSlopeThreads, 0, s>>>(out, total); count_launches(1);
      break;
    case IPX_SLOPE_CUBIC:
      if (n < 2) return IPX_EINVAL;
      slopes_cubic_kernel<T><<<grid, kSlopeThreads, 0, s>>>(a); count_launches(1);
      break;
    case IPX_SLOPE_CARDINAL: {
      if (n < 2) return IPX_EINVAL;
      T c = opts ? T(opts->c) : T(0);
      slopes_cardinal_kernel<T><<<grid, kSlopeThreads, 0, s>>>(a, T(1) - c); count_launches(1);
      break;
    }
    case IPX_SLOPE_MONOTONIC:
    case IPX_SLOPE_MONOTONIC0:
      if (n < 2) return IPX_EINVAL;
      slopes_monotonic_kernel<T><<<grid, kSlopeThreads, 0, s>>>(a, method == IPX_SLOPE_MONOTONIC0); count_launches(1);
      break;
    case IPX_SLOPE_AKIMA: {
      if (n < 2 || workspace == nullptr) return IPX_EINVAL;
      T* gmax = static_cast<T*>(workspace);
      akima_init_kernel<T><<<1, 1, 0, s>>>(gmax); count_launches(1);
      akima_max_kernel<T><<<grid, kSlopeThreads, 0, s>>>(a, gmax); count_launches(1);
      slopes_akima_kernel<T><<<grid, kSlopeThreads, 0, s>>>(a, gmax); count_launches(1);
      break;
    }
    case IPX_SLOPE_AKIMA_COMPLEX: {
      if (n < 2 || workspace == nullptr || (inner & 1) != 0) return IPX_EINVAL;
      T* gmax = static_cast<T*>(workspace);
      const int cgrid = slope_grid(total / 2);
      akima_init_kernel<T><<<1, 1, 0, s>>>(gmax); count_launches(1);
      akima_max_complex_kernel<T><<<cgrid, kSlopeThreads, 0, s>>>(a, gmax); count_launches(1);
      slopes_akima_complex_kernel<T><<<cgrid, kSlopeThreads, 0, s>>>(a, gmax); count_launches(1);
      break;
    }
    case IPX_SLOPE_CUBIC2: {
      if (n < 2) return IPX_EINVAL;
      BcArgs<T> bc;
      bc.start = opts ? opts->bc_start : IPX_BC_NOT_A_KNOT;
      bc.end = opts ? opts->bc_end : IPX_BC_NOT_A_KNOT;
      bc.start_value = opts ? static_cast<const T*>(opts->bc_start_value) : nullptr;
      bc.end_value = opts ? static_cast<const T*>(opts->bc_end_value) : nullptr;
      if (bc.start < 0 || bc.start > 2 || bc.end < 0 || bc.end > 2) return IPX_EINVAL;
      const int cgrid = slope_grid(outer * inner);
      if (n == 2 || (n == 3 && bc.start == IPX_BC_NOT_A_KNOT && bc.end == IPX_BC_NOT_A_KNOT)) {
        cubic2_small_kernel<T><<<cgrid, kSlopeThreads, 0, s>>>(a, bc); count_launches(1);
      } else {
        if (workspace == nullptr) return IPX_EINVAL;
        T* ws = static_cast<T*>(workspace);
        if (n > 4 * kChunk) {
          const int chunks = (n + kChunk - 1) / kChunk;
          cubic2_wfactor_kernel<T><<<(chunks + 127) / 128, 128, 0, s>>>(x, n, bc.start, bc.end, ws); count_launches(1);
          cubic2_wsolve_kernel<T><<<slope_grid(outer * inner * chunks), kSlopeThreads, 0, s>>>(a, bc, ws); count_launches(1);
        } else {
          cubic2_factor_kernel<T><<<1, 32, 0, s>>>(x, n, bc.start, bc.end, ws); count_launches(1);
          cubic2_solve_kernel<T><<<cgrid, kSlopeThreads, 0, s>>>(a, bc, ws); count_launches(1);
        }
      }
      break;
    }
    default:
      return IPX_EINVAL;
  }
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

// ------------------------------------------------------------------ fused slopes + pack
// One pass from f to the packed records the evaluation streams: every chain table of
// Interpolator2D/3D.__init__ (_spline.py:232-246, 380-393: fx, fy, fz from f; fxy from fx;
// fxz, fyz from fx, fy; fxyz from fxy) for the three-point stencils, without materialising
// the 2^d - 1 intermediate tables.  A warp owns a span of the last axis (one knot per lane,
// neighbours by shuffle) and marches along the axis before it with a rolling three-row
// window in registers, applying the same stencil arithmetic as the per-table kernels above,
// so the packed array is bit-identical to pack(slopes chain).  Interval widths and their
// reciprocals are computed once per CTA into shared memory.
//
// Window of an axis with n >= 3 knots: the three consecutive knots ws, ws+1, ws+2 with
// ws = clamp(k-1, 0, n-3); p = k - ws is 1 in the interior, 0 / 2 at the first / last knot,
// where the one-sided forms use knots (0, 1, 2) / (n-3, n-2, n-1).
constexpr int kFusedKnotCap = 6144;  // sum of the axis lengths that fits the shared tables

template <typename T> struct AxisWin {
  T h01, h12, r01, r12, r02;
  int p;
};

// sh: widths h[0..n-2] then reciprocals r[0..n-2] of this axis
template <typename T, int METHOD>
__device__ __forceinline__ AxisWin<T> axis_window(const T* __restrict__ x, const T* sh, int n,
                                                  int k, int& ws) {
  ws = min(max(k - 1, 0), n - 3);
  AxisWin<T> w;
  w.h01 = sh[ws];
  w.h12 = sh[ws + 1];
  w.r01 = sh[n - 1 + ws];
  w.r12 = sh[n + ws];
  w.r02 = T(0);
  if constexpr (METHOD == IPX_SLOPE_CARDINAL) w.r02 = safe_recip(__ldg(x + ws + 2) - __ldg(x + ws));
  w.p = k - ws;
  return w;
}

template <typename T> __device__ __forceinline__ T sel3(int p, T a0, T a1, T a2) {
  return p == 0 ? a0 : (p == 1 ? a1 : a2);
}

template <typename T, int METHOD>
__device__ __forceinline__ T stencil3(const AxisWin<T>& w, T g0, T g1, T g2, T one_minus_c) {
  if constexpr (METHOD == IPX_SLOPE_CUBIC) {
    T s01 = mul_rn(w.r01, g1 - g0), s12 = mul_rn(w.r12, g2 - g1);
    return w.p == 0 ? s01 : (w.p == 2 ? s12 : T(0.5) * add_rn(s01, s12));
  } else if constexpr (METHOD == IPX_SLOPE_CARDINAL) {
    T v = w.p == 0 ? (g1 - g0) * w.r01 : (w.p == 2 ? (g2 - g1) * w.r12 : w.r02 * (g2 - g0));
    return one_minus_c * v;
  } else {
    T s01 = mul_rn(w.r01, g1 - g0), s12 = mul_rn(w.r12, g2 - g1);
    if (w.p != 1) {
      if constexpr (METHOD == IPX_SLOPE_MONOTONIC0) return T(0);
      T hi0 = w.p == 0 ? w.r01 : w.r12, hi1 = w.p == 0 ? w.r12 : w.r01;
      T m0 = w.p == 0 ? s01 : s12, m1 = w.p == 0 ? s12 : s01;
      return pchip_edge(T(1) / hi0, T(1) / hi1, m0, m1);
    }
    T h0 = w.h01, h1 = w.h12, m0 = s01, m1 = s12;
    bool cond = (sign_of(m1) != sign_of(m0)) || (m1 == T(0)) || (m0 == T(0));
    T w1 = T(2) * h1 + h0;
    T w2 = h1 + T(2) * h0;
    T whmean = safediv(safediv(w1, m0) + safediv(w2, m1), w1 + w2);
    return cond ? T(0) : safediv(T(1), whmean);
  }
}

// The same stencils in centred form, for the axis that runs across the lanes of a warp:
// g0 is this lane's value at knot k, its neighbours k-1 / k+1 come from the adjacent lanes.
// h / r: widths and reciprocal widths of this axis in shared memory.  Must be called by all
// 32 lanes (shuffles); the result is meaningful on lanes whose neighbours hold k-1 and k+1.
template <typename T, int METHOD>
__device__ __forceinline__ T stencil_lanes(const T* __restrict__ x, const T* h, const T* r, int n,
                                           int k, T g0, T one_minus_c) {
  const unsigned full = 0xffffffffu;
  const T gm = __shfl_up_sync(full, g0, 1), gp = __shfl_down_sync(full, g0, 1);
  const int km = min(max(k - 1, 0), n - 2), kp = min(max(k, 0), n - 2);
  const bool first = k <= 0, last = k >= n - 1;
  if constexpr (METHOD == IPX_SLOPE_CARDINAL) {
    const int lo = max(k - 1, 0), hi = min(k + 1, n - 1);
    T r2 = safe_recip(__ldg(x + hi) - __ldg(x + lo));
    T v = first ? (gp - g0) * r[kp] : (last ? (g0 - gm) * r[km] : r2 * (gp - gm));
    return one_minus_c * v;
  } else {
    const T sm = mul_rn(r[km], g0 - gm), sp = mul_rn(r[kp], gp - g0);
    if constexpr (METHOD == IPX_SLOPE_CUBIC) {
      return first ? sp : (last ? sm : T(0.5) * add_rn(sm, sp));
    } else {
      T edge = T(0);
      if constexpr (METHOD == IPX_SLOPE_MONOTONIC) {
        // one-sided ends use the next / previous secant as well (_fd_derivs.py:365-386)
        const T sp_next = __shfl_down_sync(full, sp, 1), sm_prev = __shfl_up_sync(full, sm, 1);
        if (first || last) {
          T hi0 = first ? r[kp] : r[km];
          T hi1 = first ? r[min(kp + 1, n - 2)] : r[max(km - 1, 0)];
          edge = pchip_edge(T(1) / hi0, T(1) / hi1, first ? sp : sm, first ? sp_next : sm_prev);
        }
      }
      if (first || last) return edge;
      T h0 = h[km], h1 = h[kp], m0 = sm, m1 = sp;
      bool cond = (sign_of(m1) != sign_of(m0)) || (m1 == T(0)) || (m0 == T(0));
      T w1 = T(2) * h1 + h0;
      T w2 = h1 + T(2) * h0;
      T whmean = safediv(safediv(w1, m0) + safediv(w2, m1), w1 + w2);
      return cond ? T(0) : safediv(T(1), whmean);
    }
  }
}

template <typename T> struct FusedArgs {
  const T* __restrict__ x[3];
  const T* __restrict__ f;
  T* __restrict__ out;
  int32_t batch;
  int32_t n[3];
  int32_t wide;  // out is 32-byte aligned: 256-bit stores
  T one_minus_c;
};

// one record: 2^d values, 16-byte aligned at least
__device__ __forceinline__ void store_record4(double* o, bool wide, double a, double b, double c, double d) {
  if (wide) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(o), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
  } else {
    *reinterpret_cast<double2*>(o) = make_double2(a, b);
    *reinterpret_cast<double2*>(o + 2) = make_double2(c, d);
  }
}
__device__ __forceinline__ void store_record4(float* o, bool, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(o) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store_record8(double* o, bool wide, const double (&v)[8]) {
  store_record4(o, wide, v[0], v[1], v[2], v[3]);
  store_record4(o + 4, wide, v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store_record8(float* o, bool wide, const float (&v)[8]) {
  if (wide) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "f"(v[0]), "f"(v[1]),
                 "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
  } else {
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// widths and reciprocal widths of every axis into shared memory; returns per-axis offsets
template <typename T, int ND>
__device__ __forceinline__ void stage_widths(const FusedArgs<T>& a, T* sh, int* off) {
  int o = 0;
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    off[ax] = o;
    const int n = a.n[ax];
    for (int k = threadIdx.x; k < n - 1; k += blockDim.x) {
      T h = __ldg(a.x[ax] + k + 1) - __ldg(a.x[ax] + k);
      sh[o + k] = h;
      sh[o + n - 1 + k] = safe_recip(h);
    }
    o += 2 * (n - 1);
  }
  __syncthreads();
}

// A warp covers kLaneSpan consecutive knots of the last axis plus one halo knot on each side
// (lane 0 and lane 31), so every neighbour a centred stencil needs is one shuffle away.  The
// last chunk of a row is shifted back to end at the last knot (its overlap with the previous
// chunk rewrites the same values), which keeps both one-sided ends at least two lanes inside.
constexpr int kLaneSpan = 30;

// window of knot k from the shared width table of an axis, given its first knot ws
template <typename T, int METHOD>
__device__ __forceinline__ AxisWin<T> window_at(const T* __restrict__ x, const T* sh, int n, int k,
                                                int ws) {
  AxisWin<T> w;
  w.h01 = sh[ws];
  w.h12 = sh[ws + 1];
  w.r01 = sh[n - 1 + ws];
  w.r12 = sh[n + ws];
  w.r02 = T(0);
  if constexpr (METHOD == IPX_SLOPE_CARDINAL) w.r02 = safe_recip(__ldg(x + ws + 2) - __ldg(x + ws));
  w.p = k - ws;
  return w;
}

constexpr int kRowSpan = 32;  // rows a warp marches over with its rolling window

// 2-D: a warp owns kLaneSpan knots of y (lanes) and marches over kRowSpan consecutive x rows,
// keeping the three-row x window of f in registers: per node 1 load, 1 x stencil, 2 y stencils.
template <typename T, int METHOD>
__global__ void __launch_bounds__(kSlopeThreads) slopes_pack2d_kernel(const FusedArgs<T> a) {
  extern __shared__ __align__(16) unsigned char fused_smem[];
  T* sh = reinterpret_cast<T*>(fused_smem);
  int off[2];
  stage_widths<T, 2>(a, sh, off);
  const int nx = a.n[0], ny = a.n[1], B = a.batch;
  const int lane = threadIdx.x & 31;
  const int64_t inner = (int64_t)ny * B;
  const T omc = a.one_minus_c;
  const bool wide = a.wide != 0;
  const T* hy = sh + off[1];
  const T* ry = hy + (ny - 1);
  const unsigned chunks = (ny + kLaneSpan - 1) / kLaneSpan;
  const unsigned xsegs = (nx + kRowSpan - 1) / kRowSpan;
  const unsigned items = chunks * xsegs;
  const unsigned warps = gridDim.x * (blockDim.x >> 5);
  for (unsigned item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < items;
       item += warps) {
    const int xs = (int)(item / chunks), ch = (int)(item - (unsigned)xs * chunks);
    const int jbase = min(ch * kLaneSpan, max(ny - kLaneSpan, 0));
    const int j = jbase - 1 + lane, jc = min(max(j, 0), ny - 1);
    const bool valid = lane >= 1 && lane <= kLaneSpan && j < ny;
    const int ibeg = xs * kRowSpan, iend = min(nx, ibeg + kRowSpan);
    for (int b = 0; b < B; ++b) {
      const T* col = a.f + (int64_t)jc * B + b;  // (0, jc, b)
      T* orow = a.out + ((int64_t)ibeg * inner + (int64_t)jc * B + b) * 4;
      int ws = min(max(ibeg - 1, 0), nx - 3);
      T v0 = __ldg(col + (int64_t)ws * inner), v1 = __ldg(col + (int64_t)(ws + 1) * inner),
        v2 = __ldg(col + (int64_t)(ws + 2) * inner);
      T pv = __ldg(col + (int64_t)min(ws + 3, nx - 1) * inner);  // enters at the next shift
      for (int i = ibeg; i < iend; ++i) {
        const int wsn = min(max(i - 1, 0), nx - 3);
        if (wsn != ws) {  // warp-uniform
          ws = wsn;
          v0 = v1; v1 = v2; v2 = pv;
          pv = __ldg(col + (int64_t)min(ws + 3, nx - 1) * inner);
        }
        const AxisWin<T> wx = window_at<T, METHOD>(a.x[0], sh + off[0], nx, i, ws);
        const T fx = stencil3<T, METHOD>(wx, v0, v1, v2, omc);
        const T f0 = sel3(wx.p, v0, v1, v2);
        const T fy = stencil_lanes<T, METHOD>(a.x[1], hy, ry, ny, j, f0, omc);
        const T fxy = stencil_lanes<T, METHOD>(a.x[1], hy, ry, ny, j, fx, omc);
        if (valid) store_record4(orow, wide, f0, fy, fx, fxy);
        orow += inner * 4;
      }
    }
  }
}

// 3-D: a warp owns kLaneSpan knots of z (lanes) and marches over kRowSpan consecutive y rows
// of one x plane, keeping the three-row y window of (f, fx) in registers: per node 3 loads
// (the x window of the row that enters), 1 x stencil, 2 y stencils, 4 z stencils.
template <typename T, int METHOD>
__global__ void __launch_bounds__(kSlopeThreads, 3) slopes_pack3d_kernel(const FusedArgs<T> a) {
  extern __shared__ __align__(16) unsigned char fused_smem[];
  T* sh = reinterpret_cast<T*>(fused_smem);
  int off[3];
  stage_widths<T, 3>(a, sh, off);
  const int nx = a.n[0], ny = a.n[1], nz = a.n[2], B = a.batch;
  const int lane = threadIdx.x & 31;
  const int64_t sy = (int64_t)nz * B, sx = (int64_t)ny * sy;
  const T omc = a.one_minus_c;
  const bool wide = a.wide != 0;
  const T* hz = sh + off[2];
  const T* rz = hz + (nz - 1);
  const unsigned chunks = (nz + kLaneSpan - 1) / kLaneSpan;
  const unsigned ysegs = (ny + kRowSpan - 1) / kRowSpan;
  const unsigned per_plane = chunks * ysegs;
  const unsigned items = (unsigned)nx * per_plane;
  const unsigned warps = gridDim.x * (blockDim.x >> 5);
  for (unsigned item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); item < items;
       item += warps) {
    const int i = (int)(item / per_plane);
    const unsigned rem = item - (unsigned)i * per_plane;
    const int ys = (int)(rem / chunks), ch = (int)(rem - (unsigned)ys * chunks);
    const int kbase = min(ch * kLaneSpan, max(nz - kLaneSpan, 0));
    const int k = kbase - 1 + lane, kc = min(max(k, 0), nz - 1);
    const bool valid = lane >= 1 && lane <= kLaneSpan && k < nz;
    const int jbeg = ys * kRowSpan, jend = min(ny, jbeg + kRowSpan);
    int i0;
    const AxisWin<T> wx = axis_window<T, METHOD>(a.x[0], sh + off[0], nx, i, i0);
    for (int b = 0; b < B; ++b) {
      const T* plane = a.f + (int64_t)i0 * sx + (int64_t)kc * B + b;  // (i0, 0, kc, b)
      T* orow = a.out + ((((int64_t)i * ny + jbeg) * nz + kc) * B + b) * 8;
      int ws = min(max(jbeg - 1, 0), ny - 3);
      T gx[3], fr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const T* col = plane + (int64_t)(ws + c) * sy;
        T v0 = __ldg(col), v1 = __ldg(col + sx), v2 = __ldg(col + 2 * sx);
        gx[c] = stencil3<T, METHOD>(wx, v0, v1, v2, omc);  // fx at (i, ws + c, k)
        fr[c] = sel3(wx.p, v0, v1, v2);                    // f  at (i, ws + c, k)
      }
      // raw x window of the row that enters at the next shift, loaded one step ahead
      const T* ncol = plane + (int64_t)min(ws + 3, ny - 1) * sy;
      T p0 = __ldg(ncol), p1 = __ldg(ncol + sx), p2 = __ldg(ncol + 2 * sx);
      for (int j = jbeg; j < jend; ++j) {
        const int wsn = min(max(j - 1, 0), ny - 3);
        if (wsn != ws) {  // warp-uniform
          ws = wsn;
          gx[0] = gx[1]; gx[1] = gx[2];
          fr[0] = fr[1]; fr[1] = fr[2];
          gx[2] = stencil3<T, METHOD>(wx, p0, p1, p2, omc);
          fr[2] = sel3(wx.p, p0, p1, p2);
          ncol = plane + (int64_t)min(ws + 3, ny - 1) * sy;
          p0 = __ldg(ncol); p1 = __ldg(ncol + sx); p2 = __ldg(ncol + 2 * sx);
        }
        const AxisWin<T> wy = window_at<T, METHOD>(a.x[1], sh + off[1], ny, j, ws);
        T rec[8];  // record order: last axis fastest
        rec[0] = sel3(wy.p, fr[0], fr[1], fr[2]);
        rec[4] = sel3(wy.p, gx[0], gx[1], gx[2]);
        rec[2] = stencil3<T, METHOD>(wy, fr[0], fr[1], fr[2], omc);
        rec[6] = stencil3<T, METHOD>(wy, gx[0], gx[1], gx[2], omc);
#pragma unroll
        for (int m = 0; m < 8; m += 2)
          rec[m + 1] = stencil_lanes<T, METHOD>(a.x[2], hz, rz, nz, k, rec[m], omc);
        if (valid) store_record8(orow, wide, rec);
        orow += sy * 8;
      }
    }
  }
}

template <typename T, int METHOD>
static int launch_slopes_pack_m(int32_t ndim, const FusedArgs<T>& a, cudaStream_t s) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  size_t smem = 0;
  for (int ax = 0; ax < ndim; ++ax) smem += 2 * (size_t)(a.n[ax] - 1) * sizeof(T);
  const int64_t chunks = (a.n[ndim - 1] + kLaneSpan - 1) / kLaneSpan;
  const int64_t items = ndim == 2 ? chunks * ((a.n[0] + kRowSpan - 1) / kRowSpan)
                                  : (int64_t)a.n[0] * chunks * ((a.n[1] + kRowSpan - 1) / kRowSpan);
  if (items > INT32_MAX) return IPX_EUNSUPPORTED;
  const int64_t need = (items + kSlopeThreads / 32 - 1) / (kSlopeThreads / 32);
  const int64_t cap = (int64_t)sms * 8;
  const int grid = (int)(need < cap ? need : cap);
  if (ndim == 2) {
    auto k = slopes_pack2d_kernel<T, METHOD>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return IPX_ELAUNCH;
    k<<<grid, kSlopeThreads, smem, s>>>(a); count_launches(1);
  } else {
    auto k = slopes_pack3d_kernel<T, METHOD>;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return IPX_ELAUNCH;
    k<<<grid, kSlopeThreads, smem, s>>>(a); count_launches(1);
  }
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_slopes_pack(int32_t ndim, const T* const* knots, const int32_t* n, const T* f,
                              int64_t batch, int32_t method, const ipx_slope_opts* opts, T* out,
                              void* stream) {
  if (knots == nullptr || n == nullptr || batch < 0) return IPX_EINVAL;
  if (ndim < 1 || ndim > 3) return IPX_EINVAL;
  if (ndim == 1) return IPX_EUNSUPPORTED;  // two tables: slopes + pack is already one pass each
  if (method != IPX_SLOPE_CUBIC && method != IPX_SLOPE_CARDINAL && method != IPX_SLOPE_MONOTONIC &&
      method != IPX_SLOPE_MONOTONIC0)
    return (method >= 0 && method <= IPX_SLOPE_AKIMA_COMPLEX) ? IPX_EUNSUPPORTED : IPX_EINVAL;
  FusedArgs<T> a{};
  int64_t total = batch, knots_total = 0;
  for (int ax = 0; ax < ndim; ++ax) {
    if (knots[ax] == nullptr || n[ax] < 1) return IPX_EINVAL;
    if (n[ax] < 3) return IPX_EUNSUPPORTED;  // the three-knot window needs n >= 3
    a.x[ax] = knots[ax];
    a.n[ax] = n[ax];
    total *= n[ax];
    knots_total += n[ax];
  }
  if (knots_total > kFusedKnotCap || batch > (1 << 20)) return IPX_EUNSUPPORTED;
  if (ndim == 3 && (int64_t)n[0] * n[1] > INT32_MAX) return IPX_EUNSUPPORTED;
  if (total == 0) return IPX_OK;
  if (f == nullptr || out == nullptr) return IPX_EINVAL;
  if ((reinterpret_cast<uintptr_t>(out) & (2 * sizeof(T) - 1)) != 0) return IPX_EINVAL;
  a.f = f;
  a.out = out;
  a.batch = (int32_t)batch;
  a.wide = (reinterpret_cast<uintptr_t>(out) & 31) == 0;
  a.one_minus_c = T(1) - (opts != nullptr ? T(opts->c) : T(0));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (method) {
    case IPX_SLOPE_CUBIC: return launch_slopes_pack_m<T, IPX_SLOPE_CUBIC>(ndim, a, s);
    case IPX_SLOPE_CARDINAL: return launch_slopes_pack_m<T, IPX_SLOPE_CARDINAL>(ndim, a, s);
    case IPX_SLOPE_MONOTONIC: return launch_slopes_pack_m<T, IPX_SLOPE_MONOTONIC>(ndim, a, s);
    default: return launch_slopes_pack_m<T, IPX_SLOPE_MONOTONIC0>(ndim, a, s);
  }
}

// ------------------------------------------------------------------ reverse mode of approx_df
// For the methods that are LINEAR in f -- cubic (_fd_derivs.py:79-101), cardinal / catmull-rom
// (:303-324) and cubic2 (:160-300: slopes = T^-1 b(f), T from the knots and boundary kinds only,
// b linear in f plus a boundary-value constant) -- the slope map along one axis is
//   slopes = M f + const,   M = R (local) or T^-1 R (cubic2),
// with every row of R supported on at most three consecutive nodes.  Its transpose, which is what
// jax.grad of anything built on approx_df needs (tests/test_interpolate.py:542-637 differentiate
// through it), is grad_f = R^T y with y = g (local) or the solution of T^T y = g (cubic2).

// row k of R as (first node j0, three coefficients); rows of a masked (duplicate-knot) cubic2
// equation and rows that hold a boundary VALUE are zero
template <typename T>
__device__ __forceinline__ void slope_row(const T* __restrict__ x, int n, int method, T one_minus_c,
                                          int bs, int be, int k, int& j0, T (&c)[3]) {
  auto h = [&](int j) { return x[j + 1] - x[j]; };
  auto r = [&](int j) { return safe_recip(x[j + 1] - x[j]); };
  c[0] = c[1] = c[2] = T(0);
  const int jend = n >= 3 ? n - 3 : 0;     // window of the last row
  const int pa = n - 2 - jend, pb = n - 1 - jend;  // slots of f[n-2], f[n-1] in it
  if (method == IPX_SLOPE_CUBIC) {
    if (k == 0) { j0 = 0; c[0] = -r(0); c[1] = r(0); }
    else if (k == n - 1) { j0 = jend; c[pa] = -r(n - 2); c[pb] = r(n - 2); }
    else { j0 = k - 1; c[0] = T(-0.5) * r(k - 1); c[1] = T(0.5) * (r(k - 1) - r(k)); c[2] = T(0.5) * r(k); }
    return;
  }
  if (method == IPX_SLOPE_CARDINAL) {
    if (k == 0) { j0 = 0; c[0] = -one_minus_c * r(0); c[1] = one_minus_c * r(0); }
    else if (k == n - 1) { j0 = jend; c[pa] = -one_minus_c * r(n - 2); c[pb] = one_minus_c * r(n - 2); }
    else { j0 = k - 1; const T R = one_minus_c * safe_recip(x[k + 1] - x[k - 1]); c[0] = -R; c[2] = R; }
    return;
  }
  // cubic2: the right-hand side b_k(f) of cubic2_rhs
  T lo, di, up;
  bool masked;
  cubic2_row(x, n, bs, be, k, lo, di, up, masked);
  if (k == 0) {
    j0 = 0;
    if (masked) return;
    if (bs == IPX_BC_NOT_A_KNOT) {
      const T d = x[2] - x[0];
      const T A = (h(0) + T(2) * d) * h(1) * r(0) / d, B = h(0) * h(0) * r(1) / d;
      c[0] = -A; c[1] = A - B; c[2] = B;
    } else if (bs == IPX_BC_SECOND) { c[0] = T(-3); c[1] = T(3); }
    return;
  }
  if (k == n - 1) {
    j0 = jend;
    if (masked) return;
    if (be == IPX_BC_NOT_A_KNOT) {
      const T d = x[n - 1] - x[n - 3];
      const T A = h(n - 2) * h(n - 2) * r(n - 3) / d, B = (T(2) * d + h(n - 2)) * h(n - 3) * r(n - 2) / d;
      c[0] = -A; c[1] = A - B; c[2] = B;
    } else if (be == IPX_BC_SECOND) { c[pa] = T(-3); c[pb] = T(3); }
    return;
  }
  j0 = k - 1;
  if (masked) return;
  const T A = T(3) * h(k) * r(k - 1), B = T(3) * h(k - 1) * r(k);
  c[0] = -A; c[1] = A - B; c[2] = B;
}

// knot-only part of the Thomas recurrence for T^T (sub-diagonal of row k = upper_{k-1} of T,
// super-diagonal = lower_{k+1} of T); ws[0..n) = c', ws[n..2n) = den
template <typename T>
__global__ void cubic2_factor_T_kernel(const T* __restrict__ x, int n, int bs, int be, T* __restrict__ ws) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  T cprev = T(0), up_prev = T(0);
  for (int k = 0; k < n; ++k) {
    T lo, di, up, lo_next = T(0), d2, u2;
    bool masked;
    cubic2_row(x, n, bs, be, k, lo, di, up, masked);
    if (k + 1 < n) cubic2_row(x, n, bs, be, k + 1, lo_next, d2, u2, masked);
    const T den = di - up_prev * cprev;
    cprev = lo_next / den;
    ws[k] = cprev;
    ws[n + k] = den;
    ws[2 * n + k] = up_prev;  // sub-diagonal of T^T
    up_prev = up;
  }
}

// y = T^-T g, one thread per column, into `y` (same (outer, n, inner) view as g)
template <typename T>
__global__ void cubic2_solve_T_kernel(const T* __restrict__ g, T* __restrict__ y, int64_t outer, int64_t inner,
                                      int n, const T* __restrict__ ws) {
  const int64_t cols = outer * inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; col < cols; col += stride) {
    const int64_t o = col / inner, i = col - o * inner;
    const int64_t base = o * n * inner + i;
    T dprev = T(0);
    for (int k = 0; k < n; ++k) {
      dprev = (g[base + (int64_t)k * inner] - ws[2 * n + k] * dprev) / ws[n + k];
      y[base + (int64_t)k * inner] = dprev;
    }
    T ynext = T(0);
    for (int k = n - 1; k >= 0; --k) {
      ynext = y[base + (int64_t)k * inner] - ws[k] * ynext;
      y[base + (int64_t)k * inner] = ynext;
    }
  }
}

// grad_f[j] (+)= sum_k y[k] R[k][j]: one thread per element, rows j-1 .. j+1 and the two end rows
template <typename T>
__global__ void slopes_rowsT_kernel(const T* __restrict__ x, const T* __restrict__ y, T* __restrict__ gf,
                                    int64_t outer, int64_t inner, int n, int method, T one_minus_c, int bs, int be,
                                    int accumulate) {
  const int64_t total = outer * n * inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e % inner, rr = e / inner;
    const int j = (int)(rr % n);
    const int64_t base = (rr / n) * n * inner + i;
    T acc = T(0);
    auto add_row = [&](int k) {
      int j0;
      T c[3];
      slope_row(x, n, method, one_minus_c, bs, be, k, j0, c);
      const int p = j - j0;
      if (p >= 0 && p < 3) acc += c[p] * y[base + (int64_t)k * inner];
    };
    for (int k = j - 1; k <= j + 1; ++k)
      if (k >= 1 && k <= n - 2) add_row(k);
    add_row(0);
    if (n > 1) add_row(n - 1);
    gf[e] = accumulate ? gf[e] + acc : acc;
  }
}

template <typename T>
static int launch_slopes_vjp(const T* x, int32_t n, const T* g, int64_t outer, int64_t inner, int32_t method,
                             const ipx_slope_opts* opts, int32_t accumulate, T* grad_f, void* workspace,
                             void* stream) {
  if (outer < 0 || inner < 0 || n < 0) return IPX_EINVAL;
  const int64_t total = outer * n * inner;
  if (total == 0) return IPX_OK;
  if (x == nullptr || g == nullptr || grad_f == nullptr || g == grad_f) return IPX_EINVAL;
  if (n < 2) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = slope_grid(total);
  const T omc = T(1) - (opts ? T(opts->c) : T(0));
  switch (method) {
    case IPX_SLOPE_ZERO:
      if (!accumulate) { fill_zero_kernel<T><<<grid, kSlopeThreads, 0, s>>>(grad_f, total); count_launches(1); }
      break;
    case IPX_SLOPE_CUBIC:
    case IPX_SLOPE_CARDINAL:
      slopes_rowsT_kernel<T><<<grid, kSlopeThreads, 0, s>>>(x, g, grad_f, outer, inner, n, method, omc, 0, 0, accumulate);
      count_launches(1);
      break;
    case IPX_SLOPE_CUBIC2: {
      const int bs = opts ? opts->bc_start : IPX_BC_NOT_A_KNOT, be = opts ? opts->bc_end : IPX_BC_NOT_A_KNOT;
      if (bs < 0 || bs > 2 || be < 0 || be > 2) return IPX_EINVAL;
      // the two-knot and three-knot parabola cases of the forward pass are separate small systems
      if (n == 2 || (n == 3 && bs == IPX_BC_NOT_A_KNOT && be == IPX_BC_NOT_A_KNOT)) return IPX_EUNSUPPORTED;
      if (workspace == nullptr) return IPX_EINVAL;
      T* ws = static_cast<T*>(workspace);
      T* y = ws + 3 * (size_t)n;  // + one table-sized scratch for y
      cubic2_factor_T_kernel<T><<<1, 32, 0, s>>>(x, n, bs, be, ws); count_launches(1);
      cubic2_solve_T_kernel<T><<<slope_grid(outer * inner), kSlopeThreads, 0, s>>>(g, y, outer, inner, n, ws);
      count_launches(1);
      slopes_rowsT_kernel<T><<<grid, kSlopeThreads, 0, s>>>(x, y, grad_f, outer, inner, n, method, omc, bs, be, accumulate);
      count_launches(1);
      break;
    }
    default:  // monotonic, akima: not linear in f
      return (method >= 0 && method <= IPX_SLOPE_AKIMA_COMPLEX) ? IPX_EUNSUPPORTED : IPX_EINVAL;
  }
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

}  // namespace ipx

#define IPX_DEFINE_C2SOLVE(SUF, T)                                                                              \
  extern "C" int ipx_cubic2_solve_##SUF(const T* x, int32_t n, int32_t bc_start, int32_t bc_end, T* y,          \
                                        int64_t outer, int64_t inner, void* workspace, void* stream) {           \
    if (x == nullptr || y == nullptr || workspace == nullptr || n < 3 || outer < 0 || inner < 0)                \
      return IPX_EINVAL;                                                                                         \
    if (bc_start < 0 || bc_start > 2 || bc_end < 0 || bc_end > 2) return IPX_EINVAL;                            \
    if (outer * inner == 0) return IPX_OK;                                                                       \
    cudaStream_t s = static_cast<cudaStream_t>(stream);                                                          \
    T* ws = static_cast<T*>(workspace);                                                                          \
    ipx::cubic2_factor_kernel<T><<<1, 32, 0, s>>>(x, n, bc_start, bc_end, ws);                                   \
    ipx::cubic2_solve_rhs_kernel<T><<<ipx::slope_grid(outer * inner), ipx::kSlopeThreads, 0, s>>>(               \
        x, y, outer, inner, n, bc_start, bc_end, ws);                                                            \
    ipx::count_launches(2);                                                                                      \
    return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;                                          \
  }
IPX_DEFINE_C2SOLVE(f32, float)
IPX_DEFINE_C2SOLVE(f64, double)

extern "C" size_t ipx_slopes_vjp_workspace_bytes(int32_t method, int32_t n, int64_t outer, int64_t inner,
                                                 int32_t elem_size) {
  if (method != IPX_SLOPE_CUBIC2 || n <= 0 || outer <= 0 || inner <= 0) return 16;
  return ((size_t)3 * (size_t)n + (size_t)outer * (size_t)n * (size_t)inner) * (size_t)elem_size + 16;
}
extern "C" int ipx_slopes_vjp_f32(const float* x, int32_t n, const float* g, int64_t outer, int64_t inner,
                                  int32_t method, const ipx_slope_opts* opts, int32_t accumulate,
                                  float* grad_f, void* workspace, void* stream) {
  return ipx::launch_slopes_vjp<float>(x, n, g, outer, inner, method, opts, accumulate, grad_f, workspace, stream);
}
extern "C" int ipx_slopes_vjp_f64(const double* x, int32_t n, const double* g, int64_t outer, int64_t inner,
                                  int32_t method, const ipx_slope_opts* opts, int32_t accumulate,
                                  double* grad_f, void* workspace, void* stream) {
  return ipx::launch_slopes_vjp<double>(x, n, g, outer, inner, method, opts, accumulate, grad_f, workspace, stream);
}

extern "C" int ipx_slopes_pack_f32(int32_t ndim, const float* const* knots, const int32_t* n,
                                   const float* f, int64_t batch, int32_t method,
                                   const ipx_slope_opts* opts, float* out, void* stream) {
  return ipx::launch_slopes_pack<float>(ndim, knots, n, f, batch, method, opts, out, stream);
}
extern "C" int ipx_slopes_pack_f64(int32_t ndim, const double* const* knots, const int32_t* n,
                                   const double* f, int64_t batch, int32_t method,
                                   const ipx_slope_opts* opts, double* out, void* stream) {
  return ipx::launch_slopes_pack<double>(ndim, knots, n, f, batch, method, opts, out, stream);
}

extern "C" size_t ipx_slopes_workspace_bytes(int32_t method, int32_t n, int32_t elem_size) {
  if (method == IPX_SLOPE_AKIMA || method == IPX_SLOPE_AKIMA_COMPLEX) return 16;
  if (method == IPX_SLOPE_CUBIC2) return (size_t)2 * (size_t)(n > 0 ? n : 0) * (size_t)elem_size + 16;
  return 0;
}
extern "C" int ipx_slopes_f32(const float* x, int32_t n, const float* f, int64_t outer,
                              int64_t inner, int32_t method, const ipx_slope_opts* opts,
                              float* out, void* workspace, void* stream) {
  return ipx::launch_slopes<float>(x, n, f, outer, inner, method, opts, out, workspace, stream);
}
extern "C" int ipx_slopes_f64(const double* x, int32_t n, const double* f, int64_t outer,
                              int64_t inner, int32_t method, const ipx_slope_opts* opts,
                              double* out, void* workspace, void* stream) {
  return ipx::launch_slopes<double>(x, n, f, outer, inner, method, opts, out, workspace, stream);
}
