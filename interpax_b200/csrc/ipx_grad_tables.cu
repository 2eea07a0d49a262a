// ipx_grad_tables.cu -- reverse mode with respect to the KNOTS for the table form of the cubic
// evaluation and for the slope precompute (the methods ipx_grad.cu does not
// cover: cubic2, a global tridiagonal solve, and monotonic, non-linear in f).
//
// The reference's TestAD (tests/test_interpolate.py:542-637) takes jax.jacrev of
//     xp -> interp1d(x, xp, fp, method)        method in cubic, cubic2, cardinal, monotonic
//     xp -> interp{2,3}d(..., xp, ..., method)  method in cubic, cubic2, cardinal
// JAX traces two things that depend on the knots:
//   (1) the evaluation itself with the tables held fixed (_spline.py:576-589, 770-800, 1055-1099):
//       the local coordinate t = (q - x[i-1]) / dx and the dx that scales the slope tables;
//   (2) every link of the slope chain (_fd_derivs.py), whose stencil coefficients are functions
//       of the knot spacing.
// (1) is ipx_vjp_knots_tables_*: per query, the Hermite weights of every axis in dual numbers
// seeded on the two knots of the cell; (2) is ipx_slopes_knots_vjp_*: per table row, the slope (or,
// for cubic2, the residual b(x, f) - T(x) s of the tridiagonal system at the computed slopes s, with
// the cotangent pulled through T^-T first) in dual numbers seeded on the row's three knots -- and,
// for monotonic, on its three table values as well, because that method is not linear in f.
// The host layer (api.vjp_knots) chains them in the reverse of the reference's chain order.
#include <cuda_runtime.h>

#include "ipx_device.cuh"
#include "ipx_internal.cuh"

namespace ipx {

constexpr int kGtThreads = 128;

template <typename T, int N> struct DualN {
  T v, d[N];
};
template <typename T, int N> __device__ __forceinline__ DualN<T, N> dn_const(T v) {
  DualN<T, N> r;
  r.v = v;
#pragma unroll
  for (int k = 0; k < N; ++k) r.d[k] = T(0);
  return r;
}
template <typename T, int N> __device__ __forceinline__ DualN<T, N> dn_seed(T v, int k) {
  DualN<T, N> r = dn_const<T, N>(v);
  r.d[k] = T(1);
  return r;
}
template <typename T, int N> __device__ __forceinline__ DualN<T, N> operator+(const DualN<T, N>& a, const DualN<T, N>& b) {
  DualN<T, N> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <typename T, int N> __device__ __forceinline__ DualN<T, N> operator-(const DualN<T, N>& a, const DualN<T, N>& b) {
  DualN<T, N> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <typename T, int N> __device__ __forceinline__ DualN<T, N> operator*(const DualN<T, N>& a, const DualN<T, N>& b) {
  DualN<T, N> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < N; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
  return r;
}
template <typename T, int N> __device__ __forceinline__ DualN<T, N> operator*(T s, const DualN<T, N>& a) {
  DualN<T, N> r;
  r.v = s * a.v;
#pragma unroll
  for (int k = 0; k < N; ++k) r.d[k] = s * a.d[k];
  return r;
}
template <typename T, int N> __device__ __forceinline__ DualN<T, N> operator/(const DualN<T, N>& a, const DualN<T, N>& b) {
  DualN<T, N> r;
  const T inv = T(1) / b.v;
  r.v = a.v * inv;
#pragma unroll
  for (int k = 0; k < N; ++k) r.d[k] = (a.d[k] - r.v * b.d[k]) * inv;
  return r;
}
// where(d == 0, 0, 1 / d): a constant on the guarded branch
template <typename T, int N> __device__ __forceinline__ DualN<T, N> dn_recip(const DualN<T, N>& a) {
  if (a.v == T(0)) return dn_const<T, N>(T(0));
  return dn_const<T, N>(T(1)) / a;
}
// safediv(a, b) of _fd_derivs.py:427-443 (fill 0, threshold 0)
template <typename T, int N> __device__ __forceinline__ DualN<T, N> dn_safediv(const DualN<T, N>& a, const DualN<T, N>& b) {
  if (t_abs(b.v) <= T(0)) return dn_const<T, N>(T(0));
  return a / b;
}
template <typename T> __device__ __forceinline__ T gt_sign(T v) {
  return v > T(0) ? T(1) : (v < T(0) ? T(-1) : (v == T(0) ? T(0) : v));
}
__device__ __forceinline__ void gt_atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void gt_atomic_add(double* p, double v) { atomicAdd(p, v); }

// --------------------------------------------------------------------------- (1) evaluation
template <typename T, int ND> struct KtArgs {
  AxisDesc<T> ax[ND];
  const T* __restrict__ tab[1 << ND];  // reference stacking order, scalar-valued
  const T* __restrict__ cot;           // reverse mode: (nq)
  T* __restrict__ grad_knots[ND];      // reverse mode: nk entries each (lookup knots), or NULL
  const T* __restrict__ knots_dot[ND]; // forward mode: nk entries each, or NULL
  T* __restrict__ tangent;             // forward mode: (nq), written
  int32_t jvp;
  int64_t nq;
  ipx_params hp;
  const ipx_params* dp;
};

__device__ __forceinline__ int gt_table_mask(int ndim, int ff) {
  if (ndim < 3) return ff;
  const int d3[8] = {0, 1, 2, 4, 3, 5, 6, 7};  // f, fx, fy, fz, fxy, fxz, fyz, fxyz
  return d3[ff];
}

template <typename T, int ND>
__global__ void __launch_bounds__(kGtThreads) vjp_knots_tables_kernel(const KtArgs<T, ND> a) {
  __shared__ AxisConst<T> consts[ND];
  __shared__ ipx_params params_s;
  if (threadIdx.x < ND) axis_consts(a.ax[threadIdx.x], consts[threadIdx.x]);
  if (threadIdx.x == 32) params_s = a.dp ? *a.dp : a.hp;
  __syncthreads();
  const ipx_params& P = params_s;
  using D2 = DualN<T, 2>;
  const int64_t stride = (int64_t)gridDim.x * kGtThreads;
  for (int64_t qi = (int64_t)blockIdx.x * kGtThreads + threadIdx.x; qi < a.nq; qi += stride) {
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
    if (filled) {
      if (a.jvp) a.tangent[qi] = T(0);
      continue;
    }
    // hermite_weights in dual numbers over (x0, x1) of the cell
    D2 w[ND][4];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      const D2 x0 = dn_seed<T, 2>(L[ax].x0, 0), x1 = dn_seed<T, 2>(L[ax].x1, 1);
      const D2 dx = x1 - x0;
      const D2 dxi = dn_recip(dx);
      const D2 t = (dn_const<T, 2>(L[ax].q) - x0) * dxi;
      const int order = P.axis[ax].derivative;
      D2 tt0, tt1, tt2, tt3;
      const D2 zero = dn_const<T, 2>(T(0));
      if (order <= 0) {
        tt0 = dn_const<T, 2>(T(1)); tt1 = t; tt2 = t * t; tt3 = tt2 * t;
      } else if (order == 1) {
        tt0 = zero; tt1 = dxi; tt2 = (T(2) * t) * dxi; tt3 = (T(3) * (t * t)) * dxi;
      } else if (order == 2) {
        const D2 s = dxi * dxi;
        tt0 = zero; tt1 = zero; tt2 = T(2) * s; tt3 = (T(6) * t) * s;
      } else if (order == 3) {
        tt0 = zero; tt1 = zero; tt2 = zero; tt3 = T(6) * (dxi * dxi * dxi);
      } else {
        tt0 = tt1 = tt2 = tt3 = zero;
      }
      w[ax][0] = tt0 - T(3) * tt2 + T(2) * tt3;
      w[ax][1] = T(3) * tt2 - T(2) * tt3;
      w[ax][2] = (tt1 - T(2) * tt2 + tt3) * dx;
      w[ax][3] = (tt3 - tt2) * dx;
    }
    T g[ND][2];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) g[ax][0] = g[ax][1] = T(0);
#pragma unroll
    for (int ff = 0; ff < (1 << ND); ++ff) {
      const int mask = gt_table_mask(ND, ff);
#pragma unroll
      for (int c = 0; c < (1 << ND); ++c) {
        int64_t node = (c & 1) ? L[0].t1 : L[0].t0;
#pragma unroll
        for (int ax = 1; ax < ND; ++ax) node = node * a.ax[ax].n + (((c >> ax) & 1) ? L[ax].t1 : L[ax].t0);
        const T val = a.tab[ff][node];
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) {
          T other = val;
#pragma unroll
          for (int bx = 0; bx < ND; ++bx)
            if (bx != ax) other *= w[bx][2 * ((mask >> bx) & 1) + ((c >> bx) & 1)].v;
          const D2& wa = w[ax][2 * ((mask >> ax) & 1) + ((c >> ax) & 1)];
          g[ax][0] += other * wa.d[0];
          g[ax][1] += other * wa.d[1];
        }
      }
    }
    if (a.jvp) {
      T tan = T(0);
#pragma unroll
      for (int ax = 0; ax < ND; ++ax)
        if (a.knots_dot[ax] != nullptr)
          tan += g[ax][0] * a.knots_dot[ax][L[ax].i - 1] + g[ax][1] * a.knots_dot[ax][L[ax].i];
      a.tangent[qi] = tan;
      continue;
    }
    const T cot = a.cot[qi];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      if (a.grad_knots[ax] == nullptr) continue;
      gt_atomic_add(a.grad_knots[ax] + (L[ax].i - 1), cot * g[ax][0]);
      gt_atomic_add(a.grad_knots[ax] + L[ax].i, cot * g[ax][1]);
    }
  }
}

template <typename T, int ND>
static int launch_kt(const T* const* q, int64_t nq, const T* const* knots, const int32_t* const* perm,
                     const int32_t* n, const T* const* tables, int32_t periodic_mask, const ipx_params* params,
                     int32_t params_on_device, const T* cot, T* const* grad_knots, const T* const* knots_dot,
                     T* tangent, void* stream) {
  const int jvp = tangent != nullptr;
  if (nq < 0 || params == nullptr || tables == nullptr) return IPX_EINVAL;
  if (nq == 0) return IPX_OK;
  if (jvp ? knots_dot == nullptr : (cot == nullptr || grad_knots == nullptr)) return IPX_EINVAL;
  KtArgs<T, ND> a;
  a.jvp = jvp;
  a.tangent = tangent;
  for (int ax = 0; ax < ND; ++ax) {
    if (q[ax] == nullptr || knots[ax] == nullptr || n[ax] < 2) return IPX_EINVAL;
    const int per = (periodic_mask >> ax) & 1;
    const int wrap_only = (periodic_mask >> (4 + ax)) & 1;
    if (per && wrap_only) return IPX_EINVAL;
    a.ax[ax].q = q[ax];
    a.ax[ax].knots = knots[ax];
    a.ax[ax].perm = per && perm != nullptr ? perm[ax] : nullptr;
    a.ax[ax].n = n[ax];
    a.ax[ax].nk = per ? n[ax] + 2 : n[ax];
    a.ax[ax].periodic = per;
    a.ax[ax].wrap = per | wrap_only;
    a.grad_knots[ax] = jvp ? nullptr : grad_knots[ax];
    a.knots_dot[ax] = jvp ? knots_dot[ax] : nullptr;
  }
  for (int k = 0; k < (1 << ND); ++k) {
    if (tables[k] == nullptr) return IPX_EINVAL;
    a.tab[k] = tables[k];
  }
  a.cot = cot;
  a.nq = nq;
  if (params_on_device) { a.dp = params; a.hp = ipx_params{}; } else { a.dp = nullptr; a.hp = *params; }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (nq + kGtThreads - 1) / kGtThreads;
  int64_t cap = (int64_t)sms * 16;
  vjp_knots_tables_kernel<T, ND><<<(int)(need < cap ? need : cap), kGtThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int kt_dispatch(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots, const int32_t* const* perm,
                       const int32_t* n, const T* const* tables, int32_t periodic_mask, const ipx_params* params,
                       int32_t pod, const T* cot, T* const* grad_knots, const T* const* knots_dot, T* tangent,
                       void* stream) {
  if (ndim < 1 || ndim > 3 || q == nullptr || knots == nullptr || n == nullptr) return IPX_EINVAL;
  if (ndim == 1)
    return launch_kt<T, 1>(q, nq, knots, perm, n, tables, periodic_mask, params, pod, cot, grad_knots, knots_dot, tangent, stream);
  if (ndim == 2)
    return launch_kt<T, 2>(q, nq, knots, perm, n, tables, periodic_mask, params, pod, cot, grad_knots, knots_dot, tangent, stream);
  return launch_kt<T, 3>(q, nq, knots, perm, n, tables, periodic_mask, params, pod, cot, grad_knots, knots_dot, tangent, stream);
}

// --------------------------------------------------------------------------- (2) slope links
template <typename T> struct SkArgs {
  const T* __restrict__ x;        // n knots
  const T* __restrict__ f;        // source table, (outer, n, inner)
  const T* __restrict__ s;        // forward slopes of this link (cubic2 only)
  const T* __restrict__ y;        // cotangent of the slopes (cubic2: already pulled through T^-T)
  T* __restrict__ grad_x;         // n entries, accumulates
  T* __restrict__ grad_f;         // shape of f, accumulates (monotonic only; else untouched)
  const T* __restrict__ v0;       // cubic2 boundary values per column, or NULL
  const T* __restrict__ v1;
  int64_t outer, inner;
  int32_t n, method, bs, be;
  T fac;
};

// One row in dual numbers: seeds 0..2 = the three knots j0..j0+2, seeds 3..5 (N == 6) = the three
// table values.  Returns the slope (local methods) or the residual of the cubic2 equation.
template <typename T, int N>
__device__ __forceinline__ DualN<T, N> sk_row_eval(const SkArgs<T>& a, const DualN<T, N> (&X)[3],
                                                   const DualN<T, N> (&F)[3], int64_t base, int64_t col, int k, int j0);

template <typename T, int N>
__device__ __forceinline__ DualN<T, N> sk_row(const SkArgs<T>& a, int64_t base, int64_t col, int k, int j0) {
  using D = DualN<T, N>;
  D X[3], F[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    X[m] = dn_seed<T, N>(a.x[j0 + m], m);
    const T fv = a.f[base + (int64_t)(j0 + m) * a.inner];
    if constexpr (N == 6) F[m] = dn_seed<T, N>(fv, 3 + m);
    else F[m] = dn_const<T, N>(fv);
  }
  return sk_row_eval<T, N>(a, X, F, base, col, k, j0);
}

template <typename T, int N>
__device__ __forceinline__ DualN<T, N> sk_row_eval(const SkArgs<T>& a, const DualN<T, N> (&X)[3],
                                                   const DualN<T, N> (&F)[3], int64_t base, int64_t col, int k, int j0) {
  using D = DualN<T, N>;
  const int n = a.n;
  const D h0 = X[1] - X[0], h1 = X[2] - X[1];
  const D s01 = dn_recip(h0) * (F[1] - F[0]), s12 = dn_recip(h1) * (F[2] - F[1]);
  const D zero = dn_const<T, N>(T(0));
  if (a.method == IPX_SLOPE_CUBIC) {
    if (k == 0) return s01;
    if (k == n - 1) return s12;
    return T(0.5) * (s01 + s12);
  }
  if (a.method == IPX_SLOPE_CARDINAL) {
    if (k == 0) return a.fac * s01;
    if (k == n - 1) return a.fac * s12;
    return a.fac * (dn_recip(X[2] - X[0]) * (F[2] - F[0]));
  }
  if (a.method == IPX_SLOPE_MONOTONIC || a.method == IPX_SLOPE_MONOTONIC0) {
    if (k == 0 || k == n - 1) {
      if (a.method == IPX_SLOPE_MONOTONIC0) return zero;
      // _fd_derivs.py:365-378: intervals (0, 1) at the start, (n-2, n-3) at the end
      const D hi0 = k == 0 ? dn_recip(h0) : dn_recip(h1), hi1 = k == 0 ? dn_recip(h1) : dn_recip(h0);
      const D m0 = k == 0 ? s01 : s12, m1 = k == 0 ? s12 : s01;
      const D e0 = dn_const<T, N>(T(1)) / hi0, e1 = dn_const<T, N>(T(1)) / hi1;
      D d = ((T(2) * e0 + e1) * m0 - e0 * m1) / (e0 + e1);
      const bool mask = gt_sign(d.v) != gt_sign(m0.v);
      const bool mask2 = (gt_sign(m0.v) != gt_sign(m1.v)) && (t_abs(d.v) > T(3) * t_abs(m0.v));
      if (mask) return zero;
      if (mask2) return T(3) * m0;
      return d;
    }
    const bool cond = (gt_sign(s12.v) != gt_sign(s01.v)) || (s12.v == T(0)) || (s01.v == T(0));
    if (cond) return zero;
    const D w1 = T(2) * h1 + h0, w2 = h1 + T(2) * h0;
    const D whmean = dn_safediv(dn_safediv(w1, s01) + dn_safediv(w2, s12), w1 + w2);
    return dn_safediv(dn_const<T, N>(T(1)), whmean);
  }
  // cubic2: residual b_k(x, f) - (T(x) s)_k at the computed slopes s (constants)
  const T S0 = a.s[base + (int64_t)j0 * a.inner], S1 = a.s[base + (int64_t)(j0 + 1) * a.inner],
          S2 = a.s[base + (int64_t)(j0 + 2) * a.inner];
  const D d = X[2] - X[0];
  if (k == 0) {
    if (a.bs == IPX_BC_NOT_A_KNOT) {
      const D diag = h1, upper = d;
      if (diag.v == T(0)) return zero;
      const D b = ((h0 + T(2) * d) * h1 * s01 + h0 * h0 * s12) / d;
      return b - (S0 * diag + S1 * upper);
    }
    if (a.bs == IPX_BC_FIRST) return zero;
    if (h0.v == T(0)) return zero;
    const T v = a.v0 ? a.v0[col] : T(0);
    const D b = (T(-0.5) * v) * (h0 * h0) + T(3) * (F[1] - F[0]);
    return b - (T(2) * S0 + S1) * h0;
  }
  if (k == n - 1) {
    if (a.be == IPX_BC_NOT_A_KNOT) {
      const D diag = h0, lower = d;
      if (diag.v == T(0)) return zero;
      const D b = (h1 * h1 * s01 + (T(2) * d + h1) * h0 * s12) / d;
      return b - (S1 * lower + S2 * diag);
    }
    if (a.be == IPX_BC_FIRST) return zero;
    if (h1.v == T(0)) return zero;
    const T v = a.v1 ? a.v1[col] : T(0);
    const D b = (T(0.5) * v) * (h1 * h1) + T(3) * (F[2] - F[1]);
    return b - (S1 + T(2) * S2) * h1;
  }
  const D diag = T(2) * (h0 + h1);
  if (diag.v == T(0)) return zero;
  const D b = T(3) * (h1 * s01 + h0 * s12);
  return b - (S0 * h1 + S1 * diag + S2 * h0);
}

template <typename T, int N>
__global__ void __launch_bounds__(kGtThreads) slopes_knots_vjp_kernel(const SkArgs<T> a) {
  const int64_t total = a.outer * a.n * a.inner;
  const int64_t stride = (int64_t)gridDim.x * kGtThreads;
  const int n = a.n;
  for (int64_t e = (int64_t)blockIdx.x * kGtThreads + threadIdx.x; e < total; e += stride) {
    const int64_t i = e % a.inner, rr = e / a.inner;
    const int k = (int)(rr % n);
    const int64_t o = rr / n;
    const int64_t base = o * n * a.inner + i;
    const T yk = a.y[e];
    if (yk == T(0)) continue;
    const int j0 = k == 0 ? 0 : (k == n - 1 ? n - 3 : k - 1);
    const DualN<T, N> r = sk_row<T, N>(a, base, o * a.inner + i, k, j0);
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      if (r.d[m] != T(0)) gt_atomic_add(a.grad_x + j0 + m, yk * r.d[m]);
      if constexpr (N == 6) {
        const T df = r.d[3 + m];
        if (df != T(0)) gt_atomic_add(a.grad_f + base + (int64_t)(j0 + m) * a.inner, yk * df);
      }
    }
  }
}

// forward mode of one link: the directional derivative of every row along (x_dot, f_dot) -- the
// tangent of the slopes for the local methods, the right-hand side  b' - T' s  of the tangent system
// for cubic2 (the caller solves T s' = rhs with ipx_cubic2_solve_*)
template <typename T> struct SkFwdArgs {
  SkArgs<T> a;
  const T* __restrict__ x_dot;  // n entries or NULL
  const T* __restrict__ f_dot;  // shape of f or NULL
  T* __restrict__ out;          // shape of f
};
template <typename T>
__global__ void __launch_bounds__(kGtThreads) slopes_jvp_rows_kernel(const SkFwdArgs<T> w) {
  const SkArgs<T>& a = w.a;
  const int64_t total = a.outer * a.n * a.inner;
  const int64_t stride = (int64_t)gridDim.x * kGtThreads;
  const int n = a.n;
  using D = DualN<T, 1>;
  for (int64_t e = (int64_t)blockIdx.x * kGtThreads + threadIdx.x; e < total; e += stride) {
    const int64_t i = e % a.inner, rr = e / a.inner;
    const int k = (int)(rr % n);
    const int64_t o = rr / n;
    const int64_t base = o * n * a.inner + i;
    const int j0 = k == 0 ? 0 : (k == n - 1 ? n - 3 : k - 1);
    D X[3], F[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      X[m].v = a.x[j0 + m];
      X[m].d[0] = w.x_dot ? w.x_dot[j0 + m] : T(0);
      const int64_t at = base + (int64_t)(j0 + m) * a.inner;
      F[m].v = a.f[at];
      F[m].d[0] = w.f_dot ? w.f_dot[at] : T(0);
    }
    w.out[e] = sk_row_eval<T, 1>(a, X, F, base, o * a.inner + i, k, j0).d[0];
  }
}

// (cubic2: the caller passes y = T^-T g, which ipx_slopes_vjp_* leaves in its workspace)
template <typename T>
static int launch_sk(const T* x, int32_t n, const T* f, const T* s, const T* y, int64_t outer, int64_t inner,
                     int32_t method, const ipx_slope_opts* opts, T* grad_x, T* grad_f, void* stream) {
  if (outer < 0 || inner < 0 || n < 0) return IPX_EINVAL;
  const int64_t total = outer * n * inner;
  if (total == 0) return IPX_OK;
  if (x == nullptr || f == nullptr || y == nullptr || grad_x == nullptr) return IPX_EINVAL;
  if (n < 3) return IPX_EUNSUPPORTED;  // rows are differentiated over a three-knot window
  SkArgs<T> a{};
  a.x = x; a.f = f; a.s = s; a.y = y; a.grad_x = grad_x; a.grad_f = grad_f;
  a.outer = outer; a.inner = inner; a.n = n; a.method = method;
  a.fac = T(1) - (opts ? T(opts->c) : T(0));
  a.bs = opts ? opts->bc_start : IPX_BC_NOT_A_KNOT;
  a.be = opts ? opts->bc_end : IPX_BC_NOT_A_KNOT;
  a.v0 = opts ? static_cast<const T*>(opts->bc_start_value) : nullptr;
  a.v1 = opts ? static_cast<const T*>(opts->bc_end_value) : nullptr;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (total + kGtThreads - 1) / kGtThreads;
  int64_t cap = (int64_t)sms * 16;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (method) {
    case IPX_SLOPE_ZERO:
      return IPX_OK;
    case IPX_SLOPE_CUBIC:
    case IPX_SLOPE_CARDINAL:
      slopes_knots_vjp_kernel<T, 3><<<grid, kGtThreads, 0, st>>>(a);
      break;
    case IPX_SLOPE_MONOTONIC:
    case IPX_SLOPE_MONOTONIC0:
      if (grad_f == nullptr) return IPX_EINVAL;
      slopes_knots_vjp_kernel<T, 6><<<grid, kGtThreads, 0, st>>>(a);
      break;
    case IPX_SLOPE_CUBIC2:
      if (s == nullptr) return IPX_EINVAL;
      if (a.bs < 0 || a.bs > 2 || a.be < 0 || a.be > 2) return IPX_EINVAL;
      if (n == 3 && a.bs == IPX_BC_NOT_A_KNOT && a.be == IPX_BC_NOT_A_KNOT) return IPX_EUNSUPPORTED;
      slopes_knots_vjp_kernel<T, 3><<<grid, kGtThreads, 0, st>>>(a);
      break;
    default:
      return (method >= 0 && method <= IPX_SLOPE_AKIMA_COMPLEX) ? IPX_EUNSUPPORTED : IPX_EINVAL;
  }
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}
}  // namespace ipx

namespace ipx {
template <typename T>
static int launch_sk_fwd(const T* x, int32_t n, const T* f, const T* s, const T* x_dot, const T* f_dot, int64_t outer,
                         int64_t inner, int32_t method, const ipx_slope_opts* opts, T* out, void* stream) {
  if (outer < 0 || inner < 0 || n < 0) return IPX_EINVAL;
  const int64_t total = outer * n * inner;
  if (total == 0) return IPX_OK;
  if (x == nullptr || f == nullptr || out == nullptr) return IPX_EINVAL;
  if (n < 3) return IPX_EUNSUPPORTED;
  SkFwdArgs<T> w{};
  SkArgs<T>& a = w.a;
  a.x = x; a.f = f; a.s = s;
  a.outer = outer; a.inner = inner; a.n = n; a.method = method;
  a.fac = T(1) - (opts ? T(opts->c) : T(0));
  a.bs = opts ? opts->bc_start : IPX_BC_NOT_A_KNOT;
  a.be = opts ? opts->bc_end : IPX_BC_NOT_A_KNOT;
  a.v0 = opts ? static_cast<const T*>(opts->bc_start_value) : nullptr;
  a.v1 = opts ? static_cast<const T*>(opts->bc_end_value) : nullptr;
  w.x_dot = x_dot; w.f_dot = f_dot; w.out = out;
  if (method == IPX_SLOPE_CUBIC2) {
    if (s == nullptr || a.bs < 0 || a.bs > 2 || a.be < 0 || a.be > 2) return IPX_EINVAL;
    if (n == 3 && a.bs == IPX_BC_NOT_A_KNOT && a.be == IPX_BC_NOT_A_KNOT) return IPX_EUNSUPPORTED;
  } else if (method != IPX_SLOPE_CUBIC && method != IPX_SLOPE_CARDINAL && method != IPX_SLOPE_MONOTONIC &&
             method != IPX_SLOPE_MONOTONIC0) {
    return (method >= 0 && method <= IPX_SLOPE_AKIMA_COMPLEX) ? IPX_EUNSUPPORTED : IPX_EINVAL;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (total + kGtThreads - 1) / kGtThreads;
  int64_t cap = (int64_t)sms * 16;
  slopes_jvp_rows_kernel<T><<<(int)(need < cap ? need : cap), kGtThreads, 0, static_cast<cudaStream_t>(stream)>>>(w);
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}
}  // namespace ipx

#define IPX_DEFINE_GT(SUF, T)                                                                                      \
  extern "C" int ipx_vjp_knots_tables_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,    \
                                            const int32_t* const* perm, const int32_t* n, const T* const* tables,  \
                                            int32_t periodic_mask, const ipx_params* params,                       \
                                            int32_t params_on_device, const T* cotangent, T* const* grad_knots,    \
                                            void* stream) {                                                        \
    return ipx::kt_dispatch<T>(ndim, q, nq, knots, perm, n, tables, periodic_mask, params, params_on_device,       \
                               cotangent, grad_knots, nullptr, nullptr, stream);                                   \
  }                                                                                                                \
  extern "C" int ipx_jvp_knots_tables_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,    \
                                            const int32_t* const* perm, const int32_t* n, const T* const* tables,  \
                                            int32_t periodic_mask, const ipx_params* params,                       \
                                            int32_t params_on_device, const T* const* knots_dot, T* tangent,       \
                                            void* stream) {                                                        \
    if (tangent == nullptr) return IPX_EINVAL;                                                                     \
    return ipx::kt_dispatch<T>(ndim, q, nq, knots, perm, n, tables, periodic_mask, params, params_on_device,       \
                               nullptr, nullptr, knots_dot, tangent, stream);                                      \
  }                                                                                                                \
  extern "C" int ipx_slopes_jvp_rows_##SUF(const T* x, int32_t n, const T* f, const T* slopes, const T* x_dot,     \
                                           const T* f_dot, int64_t outer, int64_t inner, int32_t method,           \
                                           const ipx_slope_opts* opts, T* out, void* stream) {                     \
    return ipx::launch_sk_fwd<T>(x, n, f, slopes, x_dot, f_dot, outer, inner, method, opts, out, stream);          \
  }                                                                                                                \
  extern "C" int ipx_slopes_knots_vjp_##SUF(const T* x, int32_t n, const T* f, const T* slopes, const T* y,        \
                                            int64_t outer, int64_t inner, int32_t method,                          \
                                            const ipx_slope_opts* opts, T* grad_x, T* grad_f, void* stream) {      \
    return ipx::launch_sk<T>(x, n, f, slopes, y, outer, inner, method, opts, grad_x, grad_f, stream);              \
  }

IPX_DEFINE_GT(f32, float)
IPX_DEFINE_GT(f64, double)
