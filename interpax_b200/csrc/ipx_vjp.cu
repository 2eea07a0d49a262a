// ipx_vjp.cu -- transpose of the evaluation with respect to the TABLES (reverse-mode AD rule).
//
// The evaluation is linear in the tables: out[q, b] = sum_{ff, corner} tab[ff][node(q, corner), b]
// * coef(q, ff, corner), with coef a product of per-axis weights that depend on the queries and
// knots only.  Its transpose scatters a cotangent back into the tables:
//     grad[ff][node, b] += cot[q, b] * coef(q, ff, corner).
// This is what jax.grad of interp1d/2d/3d with respect to f / fx / ... computes in the
// reference by differentiating jnp.take + the A @ F contraction (interpax/_spline.py:581-589,
// 779-800, 1070-1099); here it is the custom_vjp rule of the custom call.
// Queries replaced by an extrapolation fill contribute nothing (the fill is a constant).
// SURVEY section 8(f) row f1, tables only: the dependence of the slope tables on f (approx_df)
// and of everything on the knots is not transposed here.
#include <cuda_runtime.h>

#include "ipx_device.cuh"
#include "ipx_internal.cuh"

namespace ipx {

constexpr int kVjpThreads = 256;

template <typename T, int ND> struct VjpArgs {
  AxisDesc<T> ax[ND];
  T* __restrict__ grad[1 << ND];  // reference stacking order (f, fx, ... ); [0] only for linear/nearest
  const T* __restrict__ cot;      // (nq, batch)
  int64_t nq, batch;
  ipx_params hp;
  const ipx_params* dp;
};

__device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(double* p, double v) { atomicAdd(p, v); }

// which axes the reference's ff-th table is differentiated along (bit ax)
__device__ __forceinline__ int table_mask(int ndim, int ff) {
  if (ndim == 1) return ff;                           // f, fx
  if (ndim == 2) return ff;                           // f, fx, fy, fxy  (x = bit 0, y = bit 1)
  const int d3[8] = {0, 1, 2, 4, 3, 5, 6, 7};         // f, fx, fy, fz, fxy, fxz, fyz, fxyz
  return d3[ff];
}

template <typename T, int ND, int METHOD>
__global__ void __launch_bounds__(kVjpThreads) vjp_tables_kernel(const VjpArgs<T, ND> a) {
  __shared__ AxisConst<T> consts[ND];
  __shared__ ipx_params params_s;
  if (threadIdx.x < ND) axis_consts(a.ax[threadIdx.x], consts[threadIdx.x]);
  if (threadIdx.x == 32) params_s = a.dp ? *a.dp : a.hp;
  __syncthreads();
  const ipx_params& P = params_s;
  const int64_t stride = (int64_t)gridDim.x * kVjpThreads;
  for (int64_t qi = (int64_t)blockIdx.x * kVjpThreads + threadIdx.x; qi < a.nq; qi += stride) {
    AxisLoc<T> L[ND];
    bool filled = false;
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      locate(a.ax[ax], consts[ax], P.axis[ax], qi, L[ax]);
      if (!a.ax[ax].wrap) {
        if (!P.axis[ax].keep_lo && L[ax].below) filled = true;
        if (!P.axis[ax].keep_hi && L[ax].above) filled = true;
      }
    }
    if (filled) continue;
    auto node_of = [&](int corner) {
      int64_t idx = (corner & 1) ? L[0].t1 : L[0].t0;
#pragma unroll
      for (int ax = 1; ax < ND; ++ax) idx = idx * a.ax[ax].n + (((corner >> ax) & 1) ? L[ax].t1 : L[ax].t0);
      return idx;
    };
    if (METHOD == IPX_CUBIC) {
      T w[ND][4];
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) hermite_weights(L[ax], P.axis[ax].derivative, w[ax]);
#pragma unroll
      for (int ff = 0; ff < (1 << ND); ++ff) {
        const int mask = table_mask(ND, ff);
#pragma unroll
        for (int c = 0; c < (1 << ND); ++c) {
          T coef = T(1);
#pragma unroll
          for (int ax = 0; ax < ND; ++ax) coef *= w[ax][2 * ((mask >> ax) & 1) + ((c >> ax) & 1)];
          const int64_t node = node_of(c);
          for (int64_t b = 0; b < a.batch; ++b)
            atomic_add(a.grad[ff] + node * a.batch + b, a.cot[qi * a.batch + b] * coef);
        }
      }
    } else if (METHOD == IPX_LINEAR) {
      if (ND == 1) {
        const int order = P.axis[0].derivative;
        if (order >= 2) continue;
        const T dx = L[0].x1 - L[0].x0;
        const T dxi = safe_recip(dx);
        T c0, c1;
        if (order == 1) { c0 = -dxi; c1 = dxi; }
        else if (dx == T(0)) { c0 = T(0); c1 = T(1); }
        else { const T s = (L[0].q - L[0].x0) * dxi; c0 = T(1) - s; c1 = s; }
        for (int64_t b = 0; b < a.batch; ++b) {
          const T g = a.cot[qi * a.batch + b];
          atomic_add(a.grad[0] + (int64_t)L[0].t0 * a.batch + b, g * c0);
          atomic_add(a.grad[0] + (int64_t)L[0].t1 * a.batch + b, g * c1);
        }
      } else {
        T w[ND][2];
        T scale = T(1);
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) {
          T dxi;
          linear_weights(L[ax], P.axis[ax].derivative, w[ax], dxi);
          scale = ax == 0 ? dxi : scale * dxi;
        }
#pragma unroll
        for (int c = 0; c < (1 << ND); ++c) {
          T coef = scale;
#pragma unroll
          for (int ax = 0; ax < ND; ++ax) coef *= w[ax][(c >> ax) & 1];
          const int64_t node = node_of(c);
          for (int64_t b = 0; b < a.batch; ++b)
            atomic_add(a.grad[0] + node * a.batch + b, a.cot[qi * a.batch + b] * coef);
        }
      }
    }
  }
}

template <typename T, int ND>
static int launch_vjp(const T* const* q, int64_t nq, const T* const* knots, const int32_t* const* perm,
                      const int32_t* n, int64_t batch, int32_t method, int32_t periodic_mask,
                      const ipx_params* params, int32_t params_on_device, const T* cot, T* const* grad,
                      void* stream) {
  if (nq < 0 || batch < 1 || params == nullptr || grad == nullptr) return IPX_EINVAL;
  if (method == IPX_NEAREST) return IPX_EUNSUPPORTED;  // selection, not a smooth map of the table
  if (method != IPX_LINEAR && method != IPX_CUBIC) return IPX_EINVAL;
  if (nq == 0) return IPX_OK;
  if (cot == nullptr) return IPX_EINVAL;
  VjpArgs<T, ND> a;
  for (int ax = 0; ax < ND; ++ax) {
    if (q[ax] == nullptr || knots[ax] == nullptr || n[ax] < 2) return IPX_EINVAL;
    const int per = (periodic_mask >> ax) & 1;
    const int wrap_only = (periodic_mask >> (4 + ax)) & 1;
    if (per && wrap_only) return IPX_EINVAL;
    a.ax[ax].q = q[ax];
    a.ax[ax].knots = knots[ax];
    a.ax[ax].perm = per ? perm[ax] : nullptr;
    a.ax[ax].n = n[ax];
    a.ax[ax].nk = per ? n[ax] + 2 : n[ax];
    a.ax[ax].periodic = per;
    a.ax[ax].wrap = per | wrap_only;
  }
  const int ntab = method == IPX_CUBIC ? (1 << ND) : 1;
  for (int k = 0; k < (1 << ND); ++k) a.grad[k] = nullptr;
  for (int k = 0; k < ntab; ++k) {
    if (grad[k] == nullptr) return IPX_EINVAL;
    a.grad[k] = grad[k];
  }
  a.cot = cot;
  a.nq = nq;
  a.batch = batch;
  if (params_on_device) { a.dp = params; a.hp = ipx_params{}; } else { a.dp = nullptr; a.hp = *params; }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (nq + kVjpThreads - 1) / kVjpThreads;
  int64_t cap = (int64_t)sms * 16;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (method == IPX_CUBIC) vjp_tables_kernel<T, ND, IPX_CUBIC><<<grid, kVjpThreads, 0, s>>>(a);
  else vjp_tables_kernel<T, ND, IPX_LINEAR><<<grid, kVjpThreads, 0, s>>>(a);
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int vjp_dispatch(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,
                        const int32_t* const* perm, const int32_t* n, int64_t batch, int32_t method,
                        int32_t periodic_mask, const ipx_params* params, int32_t params_on_device,
                        const T* cot, T* const* grad, void* stream) {
  if (ndim < 1 || ndim > 3 || q == nullptr || knots == nullptr || perm == nullptr || n == nullptr)
    return IPX_EINVAL;
  if (ndim == 1) return launch_vjp<T, 1>(q, nq, knots, perm, n, batch, method, periodic_mask, params, params_on_device, cot, grad, stream);
  if (ndim == 2) return launch_vjp<T, 2>(q, nq, knots, perm, n, batch, method, periodic_mask, params, params_on_device, cot, grad, stream);
  return launch_vjp<T, 3>(q, nq, knots, perm, n, batch, method, periodic_mask, params, params_on_device, cot, grad, stream);
}

}  // namespace ipx

extern "C" int ipx_vjp_tables_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                                  const int32_t* const* perm, const int32_t* n, int64_t batch,
                                  int32_t method_class, int32_t periodic_mask, const ipx_params* params,
                                  int32_t params_on_device, const float* cotangent, float* const* grad_tables,
                                  void* stream) {
  return ipx::vjp_dispatch<float>(ndim, q, nq, knots, perm, n, batch, method_class, periodic_mask, params,
                                  params_on_device, cotangent, grad_tables, stream);
}
extern "C" int ipx_vjp_tables_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                                  const int32_t* const* perm, const int32_t* n, int64_t batch,
                                  int32_t method_class, int32_t periodic_mask, const ipx_params* params,
                                  int32_t params_on_device, const double* cotangent,
                                  double* const* grad_tables, void* stream) {
  return ipx::vjp_dispatch<double>(ndim, q, nq, knots, perm, n, batch, method_class, periodic_mask, params,
                                   params_on_device, cotangent, grad_tables, stream);
}
