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
template <typename T, bool B1>
__global__ void ppoly_eval_kernel(const T* __restrict__ c, int k, int m, int64_t batch,
                                  const T* __restrict__ x, const T* __restrict__ q, int64_t nq, int nu,
                                  int extrap, T* __restrict__ out, int32_t* __restrict__ idx_out) {
  const int n = m + 1;
  const T x0 = x[0], xe = x[m];
  const T inv_h = xe > x0 ? T(m) / (xe - x0) : T(0);
  const int64_t total = nq * batch;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int terms = k - nu;  // coefficients that survive the derivative
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t iq = B1 ? e : e / batch;
    const int64_t b = B1 ? 0 : e - iq * batch;
    T v = q[iq];
    if (extrap == 2) v = x0 + py_mod(v - x0, xe - x0);  // _ppoly.py:233-235
    T lo, hi;
    const int i = bin_index(x, n, v, x0, inv_h, lo, hi);  // :239
    const T t = v - lo;                                    // :241
    const T* cp = c + (int64_t)(i - 1) * batch + b;
    const int64_t cs = (int64_t)m * batch;
    T y = T(0);
    if (nu == 0) {
      for (int j = 0; j < terms; ++j) y = add_rn(mul_rn(y, t), __ldg(cp + j * cs));
    } else {
      for (int j = 0; j < terms; ++j)
        y = add_rn(mul_rn(y, t), mul_rn(__ldg(cp + j * cs), falling<T>(k - 1 - j, nu)));
    }
    if (extrap != 1 && (v > xe || v < x0)) y = T(CUDART_NAN);  // :250-253
    out[e] = y;
    if (idx_out != nullptr && b == 0) idx_out[iq] = i;
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
static int launch_ppoly_eval(const T* c, int32_t k, int32_t m, int64_t batch, const T* x, const T* q,
                             int64_t nq, int32_t nu, int32_t extrap, T* out, int32_t* idx_out,
                             void* stream) {
  if (k < 1 || k > kPolyMaxOrder || m < 1 || batch < 0 || nq < 0 || nu < 0 || extrap < 0 || extrap > 2)
    return IPX_EINVAL;
  if (nq == 0 || batch == 0) return IPX_OK;
  if (c == nullptr || x == nullptr || q == nullptr || out == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = poly_grid(nq * batch);
  if (batch == 1)
    ppoly_eval_kernel<T, true><<<grid, kPolyThreads, 0, s>>>(c, k, m, batch, x, q, nq, nu, extrap, out, idx_out);
  else
    ppoly_eval_kernel<T, false><<<grid, kPolyThreads, 0, s>>>(c, k, m, batch, x, q, nq, nu, extrap, out, idx_out);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_from_hermite(const T* x, int32_t n, const T* y, const T* dydx, int64_t batch, T* c,
                               void* stream) {
  if (n < 2 || batch < 0) return IPX_EINVAL;
  if (batch == 0) return IPX_OK;
  if (x == nullptr || y == nullptr || dydx == nullptr || c == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  hermite_to_ppoly_kernel<T><<<poly_grid((int64_t)(n - 1) * batch), kPolyThreads, 0, s>>>(x, n, y, dydx, batch, c);
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
  ppoly_derivative_kernel<T><<<poly_grid((int64_t)(k - nu) * per), kPolyThreads, 0, s>>>(c, k, per, nu, c2);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_ppoly_antiderivative(const T* c, int32_t k, int32_t m, int64_t batch, const T* x, T* c2,
                                       void* stream) {
  if (k < 1 || k >= kPolyMaxOrder || m < 1 || batch < 0) return IPX_EINVAL;
  if (batch == 0) return IPX_OK;
  if (c == nullptr || x == nullptr || c2 == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ppoly_polyint_kernel<T><<<poly_grid((int64_t)m * batch), kPolyThreads, 0, s>>>(c, k, m, batch, x, c2);
  ppoly_continuity_kernel<T><<<poly_grid(batch), kPolyThreads, 0, s>>>(k, m, batch, c2);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

}  // namespace ipx

#define IPX_PPOLY_DEFS(SUF, T)                                                                      \
  extern "C" int ipx_ppoly_eval_##SUF(const T* c, int32_t k, int32_t m, int64_t batch, const T* x,  \
                                      const T* q, int64_t nq, int32_t nu, int32_t extrap, T* out,   \
                                      int32_t* idx_out, void* stream) {                             \
    return ipx::launch_ppoly_eval<T>(c, k, m, batch, x, q, nq, nu, extrap, out, idx_out, stream);    \
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
