// ipx_ppoly.cu -- the piecewise-polynomial family: PPoly.__call__ and the coefficient
// transforms behind CubicHermiteSpline / PchipInterpolator / Akima1DInterpolator /
// CubicSpline, PPoly.derivative and PPoly.antiderivative.
//
// Replaces interpax/_ppoly.py:189-257 (evaluation), :259-346 (derivative, antiderivative) and
// :540-569 (Hermite data -> power-basis coefficients) of the reference.  Coefficients are
// stored as the reference stores them after its moveaxis: c[(j * m + i) * batch + b] is the
// coefficient of (x - x[i])^(k-1-j) of piece i for trailing element b.
//
// Arithmetic follows the reference's operation order (polyder's coefficient x falling
// factorial with one rounding, Horner with a separate multiply and add), written with
// non-contracting intrinsics so the result does not depend on how the compiler fuses it.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "ipx_device.cuh"
#include "ipx_internal.cuh"

namespace ipx {

constexpr int kPolyThreads = 256;
constexpr int kPolyMaxOrder = 64;  // factors p (p-1) ... stay exact in double well beyond this

static int poly_grid(int64_t total) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (total + kPolyThreads - 1) / kPolyThreads;
  int64_t cap = (int64_t)sms * 16;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// falling factorial p (p-1) ... (p-nu+1), exact in T for the orders accepted here
template <typename T> __device__ __forceinline__ T falling(int p, int nu) {
  T f = T(1);
  for (int r = 0; r < nu; ++r) f *= T(p - r);
  return f;
}

// ------------------------------------------------------------------ evaluation
// extrap: 0 = NaN outside [x[0], x[m]], 1 = extend the end pieces, 2 = periodic
//
// Coefficient (j, i, b) lives at c[j * sj + i * si + b]: the reference's layout (k, m, batch) has
// sj = m * batch, si = batch; the piece-major copy ipx_ppoly_pack_* writes for scalar-valued
// polynomials, (m, k), has sj = 1, si = k, so the k coefficients of a piece share a sector.
template <typename T> struct PolyArgs {
  const T* __restrict__ c;
  const T* __restrict__ x;
  const T* __restrict__ q;
  T* __restrict__ out;
  int32_t* __restrict__ idx_out;
  int64_t sj, si, batch, nq;
  int32_t k, m, nu, extrap;
};

template <typename T> struct PolyLoc {
  T t;
  int i;
  bool outside;
};

template <typename T>
__device__ __forceinline__ PolyLoc<T> poly_locate(const PolyArgs<T>& a, T v, T x0, T xe, T inv_h) {
  PolyLoc<T> L;
  if (a.extrap == 2) v = x0 + py_mod(v - x0, xe - x0);  // _ppoly.py:233-235
  T lo, hi;
  L.i = bin_index(a.x, a.m + 1, v, x0, inv_h, lo, hi);  // :239
  L.t = v - lo;                                          // :241
  L.outside = a.extrap != 1 && (v > xe || v < x0);       // :250-253
  return L;
}

template <typename T>
__device__ __forceinline__ T poly_horner(const PolyArgs<T>& a, const T* __restrict__ cp, T t) {
  const int terms = a.k - a.nu;  // coefficients that survive the derivative
  T y = T(0);
  if (a.nu == 0) {
    for (int j = 0; j < terms; ++j) y = add_rn(mul_rn(y, t), __ldg(cp + j * a.sj));
  } else {
    for (int j = 0; j < terms; ++j)
      y = add_rn(mul_rn(y, t), mul_rn(__ldg(cp + j * a.sj), falling<T>(a.k - 1 - j, a.nu)));
  }
  return y;
}

// scalar-valued polynomials (batch == 1): one query per thread, kPolyUnroll of them in flight
constexpr int kPolyUnroll = 4;

template <typename T>
__global__ void __launch_bounds__(kPolyThreads) ppoly_eval_scalar_kernel(const PolyArgs<T> a) {
  const T x0 = a.x[0], xe = a.x[a.m];
  const T inv_h = xe > x0 ? T(a.m) / (xe - x0) : T(0);
  const int64_t span = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < a.nq; e0 += span * kPolyUnroll) {
    T v[kPolyUnroll];
#pragma unroll
    for (int u = 0; u < kPolyUnroll; ++u) {
      const int64_t e = e0 + u * span;
      v[u] = e < a.nq ? __ldcs(a.q + e) : T(0);
    }
    PolyLoc<T> L[kPolyUnroll];
#pragma unroll
    for (int u = 0; u < kPolyUnroll; ++u) L[u] = poly_locate(a, v[u], x0, xe, inv_h);
#pragma unroll
    for (int u = 0; u < kPolyUnroll; ++u) {
      const int64_t e = e0 + u * span;
      if (e >= a.nq) continue;
      T y = poly_horner(a, a.c + (int64_t)(L[u].i - 1) * a.si, L[u].t);
      if (L[u].outside) y = T(CUDART_NAN);
      __stcs(a.out + e, y);
      if (a.idx_out != nullptr) a.idx_out[e] = L[u].i;
    }
  }
}

// vector-valued polynomials: a warp owns a query, lanes stride over the trailing elements
// (coalesced coefficient reads and result writes; one lookup per query, not per element)
template <typename T>
__global__ void __launch_bounds__(kPolyThreads) ppoly_eval_rows_kernel(const PolyArgs<T> a) {
  const T x0 = a.x[0], xe = a.x[a.m];
  const T inv_h = xe > x0 ? T(a.m) / (xe - x0) : T(0);
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t iq = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); iq < a.nq; iq += warps) {
    const PolyLoc<T> L = poly_locate(a, __ldg(a.q + iq), x0, xe, inv_h);  // warp-uniform
    const T* row = a.c + (int64_t)(L.i - 1) * a.si;
    T* orow = a.out + iq * a.batch;
    for (int64_t b = lane; b < a.batch; b += 32) {
      T y = poly_horner(a, row + b, L.t);
      if (L.outside) y = T(CUDART_NAN);
      __stcs(orow + b, y);
    }
    if (a.idx_out != nullptr && lane == 0) a.idx_out[iq] = L.i;
  }
}

// few trailing elements: one (query, element) per thread
template <typename T>
__global__ void __launch_bounds__(kPolyThreads) ppoly_eval_elem_kernel(const PolyArgs<T> a) {
  const T x0 = a.x[0], xe = a.x[a.m];
  const T inv_h = xe > x0 ? T(a.m) / (xe - x0) : T(0);
  const int64_t total = a.nq * a.batch;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int B = (int)a.batch;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t iq = e / B;
    const int b = (int)(e - iq * B);
    const PolyLoc<T> L = poly_locate(a, __ldg(a.q + iq), x0, xe, inv_h);
    T y = poly_horner(a, a.c + (int64_t)(L.i - 1) * a.si + b, L.t);
    if (L.outside) y = T(CUDART_NAN);
    a.out[e] = y;
    if (a.idx_out != nullptr && b == 0) a.idx_out[iq] = L.i;
  }
}

// (k, m) -> (m, k): the k coefficients of a piece contiguous
template <typename T>
__global__ void ppoly_pack_kernel(const T* __restrict__ c, int k, int64_t m, T* __restrict__ out) {
  const int64_t total = (int64_t)k * m;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / k;
    const int j = (int)(e - i * k);
    out[e] = c[(int64_t)j * m + i];
  }
}

// ------------------------------------------------------------------ Hermite -> power basis
// piece i: F = [y_i, y_{i+1}, d_i dx, d_{i+1} dx]; a = A_CUBIC F; c = [a3/dx^3, a2/dx^2, a1/dx, a0]
template <typename T>
__global__ void hermite_to_ppoly_kernel(const T* __restrict__ x, int n, const T* __restrict__ y,
                                        const T* __restrict__ d, int64_t batch, T* __restrict__ c) {
  const int m = n - 1;
  const int64_t total = (int64_t)m * batch;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t i = e / batch;
    const T dx = x[i + 1] - x[i];
    const T f0 = y[e], f1 = y[e + batch];
    const T g0 = mul_rn(d[e], dx), g1 = mul_rn(d[e + batch], dx);
    // rows of A_CUBIC (_coefs.py:8-15), summed left to right like the reference's matmul
    const T a0 = f0;
    const T a1 = g0;
    const T a2 = add_rn(add_rn(add_rn(mul_rn(T(-3), f0), mul_rn(T(3), f1)), mul_rn(T(-2), g0)), -g1);
    const T a3 = add_rn(add_rn(add_rn(mul_rn(T(2), f0), mul_rn(T(-2), f1)), g0), g1);
    const T dx2 = mul_rn(dx, dx), dx3 = mul_rn(dx2, dx);
    c[e] = a3 / dx3;
    c[e + total] = a2 / dx2;
    c[e + 2 * total] = a1 / dx;
    c[e + 3 * total] = a0;
  }
}

// ------------------------------------------------------------------ derivative of the coefficients
// c2[j] = c[j] * falling(k-1-j, nu), j < k - nu  (k - nu >= 1); per = m * batch
template <typename T>
__global__ void ppoly_derivative_kernel(const T* __restrict__ c, int k, int64_t per, int nu,
                                        T* __restrict__ c2) {
  const int64_t total = (int64_t)(k - nu) * per;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int j = (int)(e / per);
    c2[e] = mul_rn(c[e], falling<T>(k - 1 - j, nu));
  }
}

// ------------------------------------------------------------------ one antiderivative step
// (a) c2[j] = c[j] / (k - j), c2[k] = value of the integrated piece at its right end
template <typename T>
__global__ void ppoly_polyint_kernel(const T* __restrict__ c, int k, int m, int64_t batch,
                                     const T* __restrict__ x, T* __restrict__ c2) {
  const int64_t per = (int64_t)m * batch;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < per; e += stride) {
    const int64_t i = e / batch;
    const T dx = x[i + 1] - x[i];
    T z = T(0);
    for (int j = 0; j < k; ++j) {
      const T v = c[e + j * per] / T(k - j);
      c2[e + j * per] = v;
      z = add_rn(mul_rn(z, dx), v);
    }
    c2[e + (int64_t)k * per] = add_rn(mul_rn(z, dx), T(0));  // Horner's last step: the constant is 0
  }
}

// (b) constants: c2[k][i] = sum of the right-end values of pieces 0 .. i-1, accumulated left
// to right like cumsum (_ppoly.py:337-339); one thread per trailing element
template <typename T>
__global__ void ppoly_continuity_kernel(int k, int m, int64_t batch, T* __restrict__ c2) {
  const int64_t per = (int64_t)m * batch;
  T* row = c2 + (int64_t)k * per;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += stride) {
    T s = T(0);
    for (int i = 0; i < m; ++i) {
      const T z = row[(int64_t)i * batch + b];
      row[(int64_t)i * batch + b] = i == 0 ? T(0) : s;
      s = i == 0 ? z : add_rn(s, z);
    }
  }
}

template <typename T>
static int launch_ppoly_eval(const T* c, int32_t layout, int32_t k, int32_t m, int64_t batch, const T* x,
                             const T* q, int64_t nq, int32_t nu, int32_t extrap, T* out, int32_t* idx_out,
                             void* stream) {
  if (k < 1 || k > kPolyMaxOrder || m < 1 || batch < 0 || nq < 0 || nu < 0 || extrap < 0 || extrap > 2)
    return IPX_EINVAL;
  if (layout != IPX_SOA && layout != IPX_AOS) return IPX_EINVAL;
  if (layout == IPX_AOS && batch != 1) return IPX_EUNSUPPORTED;  // piece-major copies are scalar-valued
  if (nq == 0 || batch == 0) return IPX_OK;
  if (c == nullptr || x == nullptr || q == nullptr || out == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PolyArgs<T> a;
  a.c = c; a.x = x; a.q = q; a.out = out; a.idx_out = idx_out;
  a.batch = batch; a.nq = nq; a.k = k; a.m = m; a.nu = nu; a.extrap = extrap;
  a.sj = layout == IPX_AOS ? 1 : (int64_t)m * batch;
  a.si = layout == IPX_AOS ? k : batch;
  if (batch == 1) {
    ppoly_eval_scalar_kernel<T><<<poly_grid((nq + kPolyUnroll - 1) / kPolyUnroll), kPolyThreads, 0, s>>>(a); count_launches(1);
  } else if (batch >= 16) {
    ppoly_eval_rows_kernel<T><<<poly_grid(nq * 32), kPolyThreads, 0, s>>>(a); count_launches(1);
  } else {
    ppoly_eval_elem_kernel<T><<<poly_grid(nq * batch), kPolyThreads, 0, s>>>(a); count_launches(1);
  }
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_ppoly_pack(const T* c, int32_t k, int32_t m, T* out, void* stream) {
  if (k < 1 || m < 1) return IPX_EINVAL;
  if (c == nullptr || out == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ppoly_pack_kernel<T><<<poly_grid((int64_t)k * m), kPolyThreads, 0, s>>>(c, k, m, out); count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_from_hermite(const T* x, int32_t n, const T* y, const T* dydx, int64_t batch, T* c,
                               void* stream) {
  if (n < 2 || batch < 0) return IPX_EINVAL;
  if (batch == 0) return IPX_OK;
  if (x == nullptr || y == nullptr || dydx == nullptr || c == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  hermite_to_ppoly_kernel<T><<<poly_grid((int64_t)(n - 1) * batch), kPolyThreads, 0, s>>>(x, n, y, dydx, batch, c); count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_ppoly_derivative(const T* c, int32_t k, int32_t m, int64_t batch, int32_t nu, T* c2,
                                   void* stream) {
  if (k < 1 || k > kPolyMaxOrder || m < 1 || batch < 0 || nu < 0 || nu >= k) return IPX_EINVAL;
  if (batch == 0) return IPX_OK;
  if (c == nullptr || c2 == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t per = (int64_t)m * batch;
  ppoly_derivative_kernel<T><<<poly_grid((int64_t)(k - nu) * per), kPolyThreads, 0, s>>>(c, k, per, nu, c2); count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_ppoly_antiderivative(const T* c, int32_t k, int32_t m, int64_t batch, const T* x, T* c2,
                                       void* stream) {
  if (k < 1 || k >= kPolyMaxOrder || m < 1 || batch < 0) return IPX_EINVAL;
  if (batch == 0) return IPX_OK;
  if (c == nullptr || x == nullptr || c2 == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ppoly_polyint_kernel<T><<<poly_grid((int64_t)m * batch), kPolyThreads, 0, s>>>(c, k, m, batch, x, c2); count_launches(1);
  ppoly_continuity_kernel<T><<<poly_grid(batch), kPolyThreads, 0, s>>>(k, m, batch, c2); count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

}  // namespace ipx

#define IPX_PPOLY_DEFS(SUF, T)                                                                      \
  extern "C" int ipx_ppoly_eval_##SUF(const T* c, int32_t layout, int32_t k, int32_t m,             \
                                      int64_t batch, const T* x, const T* q, int64_t nq,            \
                                      int32_t nu, int32_t extrap, T* out, int32_t* idx_out,         \
                                      void* stream) {                                               \
    return ipx::launch_ppoly_eval<T>(c, layout, k, m, batch, x, q, nq, nu, extrap, out, idx_out,     \
                                     stream);                                                       \
  }                                                                                                 \
  extern "C" int ipx_ppoly_pack_##SUF(const T* c, int32_t k, int32_t m, T* out, void* stream) {     \
    return ipx::launch_ppoly_pack<T>(c, k, m, out, stream);                                          \
  }                                                                                                 \
  extern "C" int ipx_ppoly_from_hermite_##SUF(const T* x, int32_t n, const T* y, const T* dydx,     \
                                              int64_t batch, T* c, void* stream) {                  \
    return ipx::launch_from_hermite<T>(x, n, y, dydx, batch, c, stream);                             \
  }                                                                                                 \
  extern "C" int ipx_ppoly_derivative_##SUF(const T* c, int32_t k, int32_t m, int64_t batch,        \
                                            int32_t nu, T* c2, void* stream) {                      \
    return ipx::launch_ppoly_derivative<T>(c, k, m, batch, nu, c2, stream);                          \
  }                                                                                                 \
  extern "C" int ipx_ppoly_antiderivative_##SUF(const T* c, int32_t k, int32_t m, int64_t batch,    \
                                                const T* x, T* c2, void* stream) {                  \
    return ipx::launch_ppoly_antiderivative<T>(c, k, m, batch, x, c2, stream);                       \
  }

IPX_PPOLY_DEFS(f32, float)
IPX_PPOLY_DEFS(f64, double)
