// ipx_device.cuh -- device-side building blocks shared by the evaluation kernels.
//
// Reference behaviour restated here (paths relative to the reference repo):
//   bin lookup            interpax/_spline.py:571 (and the 8 other call sites)
//   local coordinate      interpax/_spline.py:576-579
//   monomial weights      interpax/_spline.py:1138-1151 (_get_t_der)
//   Hermite basis         interpax/_coefs.py:158-163   (A_CUBIC)
//   periodic wrap         interpax/_spline.py:1117-1122
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/ipx.h"

namespace ipx {

// ------------------------------------------------------------------------------------
// small typed helpers
// ------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T t_abs(T v);
template <> __device__ __forceinline__ float t_abs<float>(float v) { return fabsf(v); }
template <> __device__ __forceinline__ double t_abs<double>(double v) { return fabs(v); }

template <typename T> __device__ __forceinline__ T t_sqrt(T v);
template <> __device__ __forceinline__ float t_sqrt<float>(float v) { return sqrtf(v); }
template <> __device__ __forceinline__ double t_sqrt<double>(double v) { return sqrt(v); }

template <typename T> __device__ __forceinline__ T t_fmod(T a, T b);
template <> __device__ __forceinline__ float t_fmod<float>(float a, float b) { return fmodf(a, b); }
template <> __device__ __forceinline__ double t_fmod<double>(double a, double b) { return fmod(a, b); }

template <typename T> __device__ __forceinline__ int t_floor_to_int(T v);
template <> __device__ __forceinline__ int t_floor_to_int<float>(float v) { return __float2int_rd(v); }
template <> __device__ __forceinline__ int t_floor_to_int<double>(double v) { return __double2int_rd(v); }

// unfused multiply / add, used where the reference's op order must not be contracted
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// Python / jnp.remainder for a positive modulus: result in [0, P] with the sign of P.
template <typename T> __device__ __forceinline__ T py_mod(T q, T P) {
  T r = t_fmod(q, P);
  if (r != T(0) && r < T(0)) r += P;
  return r;
}

// where(dx == 0, 0, 1 / dx)
template <typename T> __device__ __forceinline__ T safe_recip(T dx) {
  return dx == T(0) ? T(0) : T(1) / dx;
}

// ------------------------------------------------------------------------------------
// interval lookup:  i = clip(#{k : x[k] <= q}, 1, n-1)   (searchsorted side="right")
//
// Guess-and-verify.  The guess assumes uniformly spaced knots (x0 and inv_h come from
// the first/last knot); it is then *verified against the real knots*, stepping one cell
// if it is off by one and falling back to a full binary search otherwise, so the result
// is exact for any non-decreasing knot vector.  NaN queries count as larger than every
// knot (XLA's total-order comparator in jnp.searchsorted), i.e. i = n-1.
// On return lo = x[i-1], hi = x[i].
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ int upper_bound_clipped(const T* __restrict__ x, int n, T q) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (x[mid] <= q) lo = mid + 1; else hi = mid;
  }
  return min(max(lo, 1), n - 1);
}

template <typename T>
__device__ __forceinline__ int bin_index(const T* __restrict__ x, int n, T q, T x_first, T inv_h,
                                         T& lo, T& hi) {
  int i;
  if (q != q) {
    i = n - 1;
    lo = x[i - 1];
    hi = x[i];
    return i;
  }
  // saturating conversion (+-inf -> INT_MAX / INT_MIN); clamp BEFORE the +1 so the
  // arithmetic cannot overflow
  i = min(max(t_floor_to_int((q - x_first) * inv_h), 0), n - 2) + 1;
  lo = x[i - 1];
  hi = x[i];
  if (q < lo) {
    if (i > 1) {
      --i;
      hi = lo;
      lo = x[i - 1];
      if (q < lo && i > 1) {
        i = upper_bound_clipped(x, i, q);  // answer is in [1, i-1]
        lo = x[i - 1];
        hi = x[i];
      }
    }
  } else if (q >= hi) {
    if (i < n - 1) {
      ++i;
      lo = hi;
      hi = x[i];
      if (q >= hi && i < n - 1) {
        i = upper_bound_clipped(x, n, q);
        lo = x[i - 1];
        hi = x[i];
      }
    }
  }
  return i;
}

// ------------------------------------------------------------------------------------
// one axis of one query, located
// ------------------------------------------------------------------------------------
template <typename T> struct AxisDesc {
  const T* __restrict__ q;           // queries along this axis
  const T* __restrict__ knots;       // nk knots (padded + sorted when periodic)
  const int32_t* __restrict__ perm;  // NULL or sorted position -> table index
  int32_t n;                         // table extent along the axis
  int32_t nk;                        // n, or n + 2 on a periodic axis
  int32_t periodic;                  // padded knot position -> table index through perm
  int32_t wrap;                      // wrap the query with the Python remainder
};

template <typename T> struct AxisLoc {
  T q, x0, x1;      // wrapped query, cell end points
  int32_t i;        // interval index (upper knot), in knot coordinates
  int32_t t0, t1;   // table indices of the lower / upper corner
  bool below, above;  // q < knots[0], q > knots[nk-1]
};

// per-block constants of one axis, computed once (see stage_axis_consts)
template <typename T> struct AxisConst {
  T x_first, x_last, inv_h;
};

template <typename T>
__device__ __forceinline__ void axis_consts(const AxisDesc<T>& A, AxisConst<T>& C) {
  C.x_first = A.knots[0];
  C.x_last = A.knots[A.nk - 1];
  T span = C.x_last - C.x_first;
  C.inv_h = span > T(0) ? T(A.nk - 1) / span : T(0);
}

template <typename T>
__device__ __forceinline__ void locate_value(const AxisDesc<T>& A, const AxisConst<T>& C,
                                             const ipx_axis_params& P, T q, AxisLoc<T>& L);
template <typename T>
__device__ __forceinline__ void locate(const AxisDesc<T>& A, const AxisConst<T>& C,
                                       const ipx_axis_params& P, int64_t qi, AxisLoc<T>& L) {
  locate_value(A, C, P, A.q[qi], L);
}
// the same for a coordinate already in a register
template <typename T>
__device__ __forceinline__ void locate_value(const AxisDesc<T>& A, const AxisConst<T>& C,
                                             const ipx_axis_params& P, T q, AxisLoc<T>& L) {
  if (A.wrap) q = py_mod(q, T(fabs(P.period)));
  L.q = q;
  L.i = bin_index(A.knots, A.nk, q, C.x_first, C.inv_h, L.x0, L.x1);
  L.below = q < C.x_first;
  L.above = q > C.x_last;
  if (A.periodic) {
    // padded position p -> sorted position (p-1) mod n -> table index perm[.]
    int s0 = L.i - 2;
    if (s0 < 0) s0 += A.n;
    int s1 = L.i - 1;
    if (s1 >= A.n) s1 -= A.n;
    L.t0 = A.perm ? A.perm[s0] : s0;
    L.t1 = A.perm ? A.perm[s1] : s1;
  } else {
    L.t0 = L.i - 1;
    L.t1 = L.i;
  }
}

// table index of padded knot position p on a periodic axis (used by 1-D nearest)
template <typename T> __device__ __forceinline__ int padded_to_table(const AxisDesc<T>& A, int p) {
  if (!A.periodic) return p;
  int s = p - 1;
  if (s < 0) s += A.n;
  if (s >= A.n) s -= A.n;
  return A.perm ? A.perm[s] : s;
}

// ------------------------------------------------------------------------------------
// cubic Hermite weights of one axis.
//   tt = n-th derivative of [1, t, t^2, t^3] times dxi^n        (_get_t_der)
//   w  = tt . A_CUBIC,  then the two slope weights carry the cell width dx
//        (the reference scales the slope tables by dx before the contraction,
//         _spline.py:583-584, 790-793, 1086-1091)
// w[0], w[1]: value at lower / upper corner;  w[2], w[3]: slope at lower / upper corner.
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void hermite_weights(const AxisLoc<T>& L, int order, T w[4]) {
  T dx = L.x1 - L.x0;
  T dxi = safe_recip(dx);
  T t = (L.q - L.x0) * dxi;
  T tt0, tt1, tt2, tt3;
  if (order <= 0) {
    tt0 = T(1); tt1 = t; tt2 = t * t; tt3 = tt2 * t;
  } else if (order == 1) {
    tt0 = T(0); tt1 = dxi; tt2 = (T(2) * t) * dxi; tt3 = (T(3) * (t * t)) * dxi;
  } else if (order == 2) {
    T s = dxi * dxi;
    tt0 = T(0); tt1 = T(0); tt2 = T(2) * s; tt3 = (T(6) * t) * s;
  } else if (order == 3) {
    tt0 = T(0); tt1 = T(0); tt2 = T(0); tt3 = T(6) * (dxi * dxi * dxi);
  } else {
    T z = dxi * T(0);
    tt0 = z; tt1 = z; tt2 = z; tt3 = z;
  }
  // columns of A_CUBIC = [[1,0,0,0],[0,0,1,0],[-3,3,-2,-1],[2,-2,1,1]]
  w[0] = tt0 - T(3) * tt2 + T(2) * tt3;
  w[1] = T(3) * tt2 - T(2) * tt3;
  w[2] = (tt1 - T(2) * tt2 + tt3) * dx;
  w[3] = (tt3 - tt2) * dx;
}

// (bi/tri)linear weights of one axis (_spline.py:745-753, 1008-1020); the common factor
// dxi is returned separately because the reference multiplies it in after the sum.
template <typename T>
__device__ __forceinline__ void linear_weights(const AxisLoc<T>& L, int order, T w[2], T& dxi) {
  dxi = safe_recip(L.x1 - L.x0);
  if (order <= 0) {
    w[0] = L.x1 - L.q; w[1] = L.q - L.x0;
  } else if (order == 1) {
    w[0] = T(-1); w[1] = T(1);
  } else {
    w[0] = T(0); w[1] = T(0);
  }
}

// ------------------------------------------------------------------------------------
// pieces shared by the warp-cooperative kernels (ipx_eval_tile.cuh, ipx_stencil.cu)
// ------------------------------------------------------------------------------------
// rare paths kept out of line (values in and out through registers only): the hot loop must
// stay small for the instruction caches
//
// Remainder of a query that is more than one period out.  fmod is a long software loop on the
// GPU; for a quotient small enough that floor(q / P) is off by at most one, the same value comes
// from one division and one fused multiply-add: q - k P is formed exactly inside the fma and
// rounded once, which is what fmod (exact) followed by the Python sign fix-up (+ P, one rounding)
// produces.  An estimate one too large shows as r < 0, one too small as r >= P -- except that a
// tiny negative q legitimately rounds up to r == P, told apart by trying k + 1.  Everything else
// (huge quotients, NaN, inf, P == 0) keeps the fmod path.
template <typename T> __device__ __forceinline__ T quotient_limit();
template <> __device__ __forceinline__ float quotient_limit<float>() { return 1048576.f; }        // 2^20
template <> __device__ __forceinline__ double quotient_limit<double>() { return 2147483648.0; }   // 2^31
template <typename T> __device__ __noinline__ T py_mod_slow(T q, T P) {
  const T k = floor(q / P);
  if (fabs(k) < quotient_limit<T>()) {
    T r = fma(-k, P, q);
    if (r < T(0)) {
      r = fma(-(k - T(1)), P, q);
    } else if (r >= P) {
      const T r2 = fma(-(k + T(1)), P, q);
      if (r2 >= T(0)) r = r2;
    }
    return r;
  }
  return py_mod(q, P);
}
template <typename T>
__device__ __noinline__ int bin_index_slow(const T* x, int n, T q, T x_first, T inv_h) {
  T lo, hi;
  return bin_index(x, n, q, x_first, inv_h, lo, hi);
}

// One-period wrap of a query: r = q + P for q in [-P, 0), q - P for q in [P, 2P) (exact, Sterbenz),
// q for q in [0, P) -- what the Python remainder gives there.  `bad` = anything else (|q| beyond one
// period, NaN, -0.0, a sum that rounds up to P): the caller repairs those with the out-of-line
// fmod.  Signs are tested on the integer pipe (q - P >= 0 <=> q >= P exactly), which leaves one
// add and one compare on the FP64 pipe that bounds the kernel.
__device__ __forceinline__ int sign_word(double v) { return __double2hiint(v); }
__device__ __forceinline__ int sign_word(float v) { return __float_as_int(v); }
template <typename T> __device__ __forceinline__ T wrap_one_period(T v, T Pd, bool& bad) {
  const bool neg = sign_word(v) < 0;
  const T s = v + (neg ? Pd : -Pd);
  const bool s_neg = sign_word(s) < 0;
  const T r = (neg || !s_neg) ? s : v;
  bad = (neg && s_neg) || !(r < Pd);
  return r;
}

// cubic Hermite weights from the local coordinate t, the cell width dx and its reciprocal
// (same formulas as hermite_weights)
template <typename T, bool ORD0>
__device__ __forceinline__ void hermite_weights_t(T t, T dx, T dxi, int order, T w[4]) {
  T tt0, tt1, tt2, tt3;
  if (ORD0 || order <= 0) {
    // order 0: the same four cubics factored, 9 operations instead of 11 (the tile kernel is bound
    // by the FP64 pipe: C4 1.94 -> 1.885 ms per 1e8 queries); a few ulp from the unfactored form
    const T t2 = t * t;
    const T u = (t - T(1)) * dx;
    w[1] = t2 * fma(T(-2), t, T(3));
    w[0] = T(1) - w[1];
    w[3] = t2 * u;
    w[2] = (t2 - t) * u;
    return;
  }
  if (ORD0 || order <= 0) {
    tt0 = T(1); tt1 = t; tt2 = t * t; tt3 = tt2 * t;
  } else if (order == 1) {
    tt0 = T(0); tt1 = dxi; tt2 = (T(2) * t) * dxi; tt3 = (T(3) * (t * t)) * dxi;
  } else if (order == 2) {
    T s = dxi * dxi;
    tt0 = T(0); tt1 = T(0); tt2 = T(2) * s; tt3 = (T(6) * t) * s;
  } else if (order == 3) {
    tt0 = T(0); tt1 = T(0); tt2 = T(0); tt3 = T(6) * (dxi * dxi * dxi);
  } else {
    T z = dxi * T(0);
    tt0 = z; tt1 = z; tt2 = z; tt3 = z;
  }
  w[0] = tt0 - T(3) * tt2 + T(2) * tt3;
  w[1] = T(3) * tt2 - T(2) * tt3;
  w[2] = (tt1 - T(2) * tt2 + tt3) * dx;
  w[3] = (tt3 - tt2) * dx;
}
template <typename T, bool ORD0>
__device__ __forceinline__ void hermite_weights_r(T q, T x0, T x1, int order, T dxi, T w[4]) {
  if (ORD0 || order <= 0) {
    // order 0, factored, with (t - 1) dx formed as q - x1 (the same quantity in one rounding):
    // 10 FP64 operations per axis, the local coordinate included
    const T t = (q - x0) * dxi;
    const T t2 = t * t;
    const T u = q - x1;  // (a zero-width cell has 1 / dx = 0, so t = 0 and both slope weights vanish as before)
    w[1] = t2 * fma(T(-2), t, T(3));
    w[0] = T(1) - w[1];
    w[3] = t2 * u;
    w[2] = (t2 - t) * u;
    return;
  }
  hermite_weights_t<T, ORD0>((q - x0) * dxi, x1 - x0, dxi, order, w);
}

// ------------------------------------------------------------------------------------
// global loads.  Tables are read-only for the lifetime of a call: use the non-coherent
// path; 128-bit where the record layout allows it.
// ------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }

__device__ __forceinline__ void ldg_rec2(const double* p, double* r) {
  double2 v = __ldg(reinterpret_cast<const double2*>(p));
  r[0] = v.x; r[1] = v.y;
}
__device__ __forceinline__ void ldg_rec2(const float* p, float* r) {
  float2 v = __ldg(reinterpret_cast<const float2*>(p));
  r[0] = v.x; r[1] = v.y;
}
__device__ __forceinline__ void ldg_rec4(const double* p, double* r) {
  ldg_rec2(p, r); ldg_rec2(p + 2, r + 2);
}
__device__ __forceinline__ void ldg_rec4(const float* p, float* r) {
  float4 v = __ldg(reinterpret_cast<const float4*>(p));
  r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
}
__device__ __forceinline__ void ldg_rec8(const double* p, double* r) {
  ldg_rec4(p, r); ldg_rec4(p + 4, r + 4);
}
__device__ __forceinline__ void ldg_rec8(const float* p, float* r) {
  ldg_rec4(p, r); ldg_rec4(p + 4, r + 4);
}

}  // namespace ipx
