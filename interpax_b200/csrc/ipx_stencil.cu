// ipx_stencil.cu -- cubic evaluation straight from `f` for the three-point slope stencils
// (method = cubic, cardinal, catmull-rom), SURVEY 8(f) row f2.
//
// What the reference does (paths in the reference repo): the slope tables fx, fy, fz, fxy, ...
// are chained applications of approx_df (interpax/_spline.py:380-393 for the classes,
// :1027-1040 for the functional form; stencils interpax/_fd_derivs.py:79-101 `_cubic1` and
// :303-324 `_cardinal`), then 2^d tables are gathered at the 2^d cell corners and contracted
// with A_CUBIC (x) ... (x) A_CUBIC (_spline.py:1070-1099).  For these two stencils every link
// of the chain is LINEAR in f and acts along ONE axis, so the whole thing collapses, per axis,
// into four tap weights on the nodes i-2 .. i+1 around the cell [i-1, i]:
//
//     slope(i-1) = aL f[i-2] + bL f[i-1] + cL f[i]      slope(i) = aR f[i-1] + bR f[i] + cR f[i+1]
//     T0 = w2 aL      T1 = w0 + w2 bL + w3 aR      T2 = w1 + w2 cL + w3 bR      T3 = w3 cR
//
// with (w0, w1, w2, w3) the cubic Hermite weights of _get_t_der . A_CUBIC (hermite_weights_r)
// and (a, b, c) the stencil coefficients of the node (one-sided at the ends of the knot vector
// the slopes were taken on).  The value is the separable contraction of the 4^d neighbourhood
// of f with Tx (x) Ty (x) Tz: the same 64 + 16 + 4 fused multiply-adds as the record form, but
// the table is 2^d times smaller (256^3 f64: 134 MB instead of 1.07 GB, about the size of L2),
// neighbouring cells share 3/4 of their operands per axis, and neither the slope chain nor the
// periodic pad of the functional form is materialised.
//
// Table: the "halo" copy of f written by ipx_halo_table_*: extent nk + 2 per axis (nk = number
// of lookup knots: n, or n + 2 on a periodic axis), entry u holds the value at lookup-knot
// position u - 1 (periodic axes: table index perm[(u - 2) mod n]; open axes: zero outside), so
// the taps of cell i are the contiguous entries i-1 .. i+2 on every axis, with no index
// arithmetic in the kernel.  The two entry points of the reference differ only in which knots
// the slopes see: Interpolator*: the un-padded knots, one-sided at table rows 0 and
// n-1, then wrapped; interpNd(..., period): the padded knots, one-sided at padded rows 0 and
// n+1.  That is a per-axis choice of coefficient table (StAxis::ring), built per CTA in the
// prologue from the knots.
//
// Kernel shapes (one persistent CTA per SM, knots + coefficients staged in shared memory):
//   Q = 1  one query per lane: sparse or incoherent streams; 4^(d-1) rows x 4 taps from L1/L2.
//   Q = 4  four CONSECUTIVE queries per lane: dense cell-sorted streams.  The four share the
//          cell column (all axes but the last) and sit within TAPS - 4 cells of each other
//          along the last axis, so each of the 4^(d-1) rows is loaded ONCE (TAPS values) and
//          applied to all four with zero-padded last-axis taps: operand delivery through the
//          LSU, which bounds a thread-per-query tricubic at 4 clk per query per SM, drops to
//          TAPS/16 of that.  A query that does not fit (other column, further away) and any
//          NaN result (a non-finite table entry under a zero tap) is redone by the
//          one-query path, so results never depend on the neighbours in the stream.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "ipx_device.cuh"
#include "ipx_halo.cuh"
#include "ipx_internal.cuh"

namespace ipx {

// elements per row of the halo table: the last axis padded to a multiple of 32 bytes
template <typename T> __host__ __device__ constexpr int64_t st_halo_pitch(int64_t ext) {
  constexpr int64_t A = 32 / (int64_t)sizeof(T);
  return (ext + A - 1) / A * A;
}

template <typename T> struct StAxis {
  const T* __restrict__ q;       // queries along this axis
  const T* __restrict__ knots;   // nk lookup knots (padded + sorted on a periodic axis)
  const T* __restrict__ sknots;  // ns knots the slope stencil is taken on
  int64_t stride;                // halo-table stride of this axis, in elements
  int32_t nk, ns;
  int32_t ring;  // slope node of lookup knot p: (p - 1) mod ns if ring (class form, periodic), else p
  int32_t wrap;  // wrap the query with the Python remainder
};

template <typename T, int ND> struct StArgs {
  StAxis<T> ax[ND];
  const T* __restrict__ tab;  // halo table
  int64_t pitch;              // elements per row of it (the padded last extent)
  T* __restrict__ out;
  int32_t* __restrict__ idx_out;
  int64_t nq;
  T fac;             // cardinal: 1 - c
  int32_t cardinal;  // 0: _cubic1, 1: _cardinal
  const int32_t* __restrict__ flag;  // NULL, or: run only if bit *flag of `accept` is set (device-side
  uint32_t accept;                   // choice of kernel shape and table layout, see launch_stencil)
  // Visiting order (ipx_eval_stencil_perm_*): query k of the launch is the caller's query order[k],
  // its coordinates come from the interleaved records of ipx_cell_order_records_* (one sector per
  // query) and its result goes to out[order[k]] -- the gather and scatter passes of a pre-sorted
  // call folded into the evaluation.  Sparse shape only.
  const uint32_t* __restrict__ order;
  const T* __restrict__ recs;
  ipx_params hp;
  const ipx_params* dp;
};

// stencil coefficients (a, c) of one lookup knot; b = -(a + c) (the three sum to zero)
template <typename T> struct alignas(2 * sizeof(T)) Coef2 {
  T a, c;
};

// per-axis constants that need device data (shared memory, written once per CTA)
template <typename T> struct StAxC {
  T x_first, inv_h, x_last, pad;
};

template <int ND> struct StLayout {
  int ks_off[ND];  // byte offsets into dynamic shared memory
  int rr_off[ND];
  int ac_off[ND];
  int total;
};

template <typename T, int ND> __host__ __device__ inline StLayout<ND> st_layout(const StArgs<T, ND>& a) {
  StLayout<ND> L;
  int off = 0;
  for (int ax = 0; ax < ND; ++ax) {
    const int col = (a.ax[ax].nk * (int)sizeof(T) + 15) & ~15;
    L.ks_off[ax] = off;
    off += col;
    L.rr_off[ax] = off;
    off += col;
    L.ac_off[ax] = off;
    off += (a.ax[ax].nk * (int)sizeof(Coef2<T>) + 15) & ~15;
  }
  L.total = off;  // a multiple of 16: per-warp scratch of the dense kernel follows
  return L;
}

// shared-memory views of one CTA
template <typename T, int ND> struct StShared {
  const T* ks[ND];          // lookup knots
  const T* rr[ND];          // rr[i]: reciprocal width of cell i = [ks[i-1], ks[i]]; sign bit set = not a "regular" cell
  const Coef2<T>* ac[ND];   // slope-stencil coefficients of the knots
  const StAxC<T>* axc;
};

// slope-stencil coefficients of node m of the ns knots xs (_fd_derivs.py:79-101, 303-324)
template <typename T>
__device__ __forceinline__ void st_node_coefs(const T* __restrict__ xs, int ns, int m, bool cardinal, T fac, T& ca,
                                              T& cc) {
  const T r_lo = m > 0 ? safe_recip(xs[m] - xs[m - 1]) : T(0);
  const T r_hi = m < ns - 1 ? safe_recip(xs[m + 1] - xs[m]) : T(0);
  if (m == 0) {
    ca = T(0); cc = cardinal ? fac * r_hi : r_hi;
  } else if (m == ns - 1) {
    ca = -(cardinal ? fac * r_lo : r_lo); cc = T(0);
  } else if (cardinal) {
    const T s = fac * safe_recip(xs[m + 1] - xs[m - 1]);
    ca = -s; cc = s;
  } else {
    ca = T(-0.5) * r_lo; cc = T(0.5) * r_hi;
  }
}

// Stage knots, reciprocal widths and stencil coefficients of every axis (all threads of the CTA;
// ends with a barrier).  A cell is REGULAR when the axis is uniformly spaced to within `tol`
// (lookup knots and slope knots, same spacing), the stencil is (f[m+1] - f[m-1]) / 2h (cubic;
// cardinal with c = 0) and both of its knots carry it -- the ends of an open or pad-first axis do,
// through the ghost entries of the halo table; the seam of a class-form periodic axis does not:
// its four taps are then fixed cubics in t (st_taps), with the spacing RATIOS of neighbouring
// cells taken as exactly 1 -- an error of at most 0.15 * tol * max|f| -- while t itself still
// comes from the cell's own knot and reciprocal width.
template <typename T, int ND>
__device__ __forceinline__ void st_stage(const StArgs<T, ND>& a, char* smem, StAxC<T>* AXC, StShared<T, ND>& S) {
  const T tol = sizeof(T) == 8 ? T(1e-12) : T(1e-5);
  if (threadIdx.x < ND) {
    const int ax = threadIdx.x;
    const StAxis<T>& A = a.ax[ax];
    const T xf = A.knots[0], xl = A.knots[A.nk - 1];
    const T span = xl - xf;
    AXC[ax].x_first = xf; AXC[ax].x_last = xl;
    AXC[ax].inv_h = span > T(0) ? T(A.nk - 1) / span : T(0);
    AXC[ax].pad = T(0);
  }
  const StLayout<ND> lay = st_layout<T, ND>(a);
  S.axc = AXC;
  __syncthreads();
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    const StAxis<T>& A = a.ax[ax];
    T* ks = reinterpret_cast<T*>(smem + lay.ks_off[ax]);
    T* rr = reinterpret_cast<T*>(smem + lay.rr_off[ax]);
    Coef2<T>* ac = reinterpret_cast<Coef2<T>*>(smem + lay.ac_off[ax]);
    S.ks[ax] = ks;
    S.rr[ax] = rr;
    S.ac[ax] = ac;
    const T inv_h = AXC[ax].inv_h;
    const bool centred = !a.cardinal || a.fac == T(1);
    bool uniform = centred;
    for (int p = threadIdx.x; p < A.nk; p += (int)blockDim.x) {
      const T xp = A.knots[p];
      ks[p] = xp;
      int m = p;
      if (A.ring) {
        m = (p - 1) % A.ns;
        if (m < 0) m += A.ns;
      }
      Coef2<T> c;
      st_node_coefs<T>(A.sknots, A.ns, m, a.cardinal != 0, a.fac, c.a, c.c);
      ac[p] = c;
      T r = T(0);
      if (p > 0) {
        const T dx = xp - A.knots[p - 1];
        r = safe_recip(dx);
        uniform = uniform && t_abs(dx * inv_h - T(1)) <= tol;
      }
      rr[p] = r;
      if (m > 0) uniform = uniform && t_abs((A.sknots[m] - A.sknots[m - 1]) * inv_h - T(1)) <= tol;
    }
    uniform = __syncthreads_and(uniform) != 0;  // (also orders the writes above before the reads below)
    for (int i = threadIdx.x; i < A.nk; i += (int)blockDim.x) {
      bool regular = uniform && i >= 1;
      if (regular && A.ring) {
        // both knots of the cell must carry the centred stencil: not the cells around the seam,
        // whose one-sided end stencils sit between wrapped table rows
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          int m = (i - d - 1) % A.ns;
          if (m < 0) m += A.ns;
          regular = regular && m > 0 && m < A.ns - 1;
        }
      }
      // (elsewhere the one-sided end stencil IS the centred one applied to the halo's ghost
      // entry 2 f[0] - f[1] resp. 2 f[n-1] - f[n-2], see halo_kernel)
      if (!regular) rr[i] = -rr[i];
    }
  }
  __syncthreads();
}

template <typename T> __device__ __forceinline__ bool st_regular(T r) { return sign_word(r) >= 0; }

template <typename T> struct Taps4 {
  T t0, t1, t2, t3;
};

// any cell, any derivative order: Hermite weights of _get_t_der . A_CUBIC composed with the
// stencil coefficients of the cell's two knots.  Out of line: the hot loops only call it for
// the cells that are not regular (one-sided ends, non-uniform knots, cardinal tension, orders >= 2).
// (Every operation below is an explicit round-to-nearest multiply, add or fused multiply-add, so
// the compiler cannot contract or reassociate it differently from one inlining context to the
// next: the out-of-line copy the sparse / dense shapes call and the inlined copy in the column
// shape's hot loop give the same bits.)
template <typename T>
__device__ __forceinline__ Taps4<T> st_taps_general_inl(const T* __restrict__ ks, const Coef2<T>* __restrict__ ac,
                                                        int i, T t, T rabs, int order) {
  const Coef2<T> L = ac[i - 1], R = ac[i];
  const T dx = add_rn(ks[i], -ks[i - 1]);
  const T dxi = rabs;
  T tt0, tt1, tt2, tt3;
  if (order <= 0) {
    tt0 = T(1); tt1 = t; tt2 = mul_rn(t, t); tt3 = mul_rn(tt2, t);
  } else if (order == 1) {
    tt0 = T(0); tt1 = dxi; tt2 = mul_rn(mul_rn(T(2), t), dxi); tt3 = mul_rn(mul_rn(T(3), mul_rn(t, t)), dxi);
  } else if (order == 2) {
    const T s2 = mul_rn(dxi, dxi);
    tt0 = T(0); tt1 = T(0); tt2 = mul_rn(T(2), s2); tt3 = mul_rn(mul_rn(T(6), t), s2);
  } else if (order == 3) {
    tt0 = T(0); tt1 = T(0); tt2 = T(0); tt3 = mul_rn(T(6), mul_rn(mul_rn(dxi, dxi), dxi));
  } else {
    const T z = mul_rn(dxi, T(0));
    tt0 = z; tt1 = z; tt2 = z; tt3 = z;
  }
  // tt . A_CUBIC (columns [[1,0,0,0],[0,0,1,0],[-3,3,-2,-1],[2,-2,1,1]]), slope weights times dx
  const T w0 = fma(T(2), tt3, fma(T(-3), tt2, tt0));
  const T w1 = fma(T(-2), tt3, mul_rn(T(3), tt2));
  const T w2 = mul_rn(add_rn(fma(T(-2), tt2, tt1), tt3), dx);
  const T w3 = mul_rn(add_rn(tt3, -tt2), dx);
  Taps4<T> o;
  o.t0 = mul_rn(w2, L.a);
  o.t1 = fma(w3, R.a, fma(w2, -add_rn(L.a, L.c), w0));
  o.t2 = fma(w3, -add_rn(R.a, R.c), fma(w2, L.c, w1));
  o.t3 = mul_rn(w3, R.c);
  return o;
}
template <typename T>
__device__ __noinline__ Taps4<T> st_taps_general_t(const T* __restrict__ ks, const Coef2<T>* __restrict__ ac, int i,
                                                   T t, T rabs, int order) {
  return st_taps_general_inl<T>(ks, ac, i, t, rabs, order);
}
// (l = ks[i-1]; one out-of-line body for every kernel shape, so the shapes agree bit for bit)
template <typename T>
__device__ __forceinline__ Taps4<T> st_taps_general(const T* __restrict__ ks, const Coef2<T>* __restrict__ ac, int i,
                                                    T l, T r, T v, int order) {
  const T rabs = t_abs(r);
  return st_taps_general_t<T>(ks, ac, i, mul_rn(add_rn(v, -l), rabs), rabs, order);
}

// regular cells only (the dense kernel's hot loop; it redoes everything else out of line)
template <typename T, bool ORD0>
__device__ __forceinline__ void st_taps_regular(T l, T r, T v, int order, T (&tp)[4]) {
  const T t = (v - l) * r;
  if (ORD0 || order <= 0) {
    const T t2 = t * t;
    tp[1] = fma(t2, fma(T(1.5), t, T(-2.5)), T(1));
    tp[3] = t2 * fma(T(0.5), t, T(-0.5));
    tp[0] = t * fma(t, fma(T(-0.5), t, T(1)), T(-0.5));
    tp[2] = t * fma(t, fma(T(-1.5), t, T(2)), T(0.5));
  } else {
    const T tr = t * r;
    tp[0] = fma(tr, fma(T(-1.5), t, T(2)), T(-0.5) * r);
    tp[1] = tr * fma(T(4.5), t, T(-5));
    tp[2] = fma(tr, fma(T(-4.5), t, T(4)), T(0.5) * r);
    tp[3] = tr * fma(T(1.5), t, T(-1));
  }
}

// The four tap weights of one axis for a query at v in cell i (l = ks[i-1], r = rr[i]).
template <typename T, bool ORD0>
__device__ __forceinline__ void st_taps(const T* __restrict__ ks, const Coef2<T>* __restrict__ ac, int i, T l, T r,
                                        T v, int order, T (&tp)[4]) {
  if (st_regular(r) && (ORD0 || order <= 1)) {
    // centred stencil on uniform knots: slope = (f[+1] - f[-1]) / 2h, so the taps are the
    // Hermite weights of _get_t_der . A_CUBIC combined with (-1/2, 0, 1/2): fixed cubics in t
    // (order 0) or their derivatives times 1/h (order 1)
    st_taps_regular<T, ORD0>(l, r, v, order, tp);
  } else {
    const Taps4<T> g = st_taps_general<T>(ks, ac, i, l, r, v, ORD0 ? 0 : order);
    tp[0] = g.t0; tp[1] = g.t1; tp[2] = g.t2; tp[3] = g.t3;
  }
}

// wrap + interval guess of one coordinate; `bad` = the caller must take the slow path.
// On return l = ks[i-1].
template <typename T>
__device__ __forceinline__ int st_locate_fast(const T* __restrict__ ks, int nk, const StAxC<T>& C, bool wrap, T Pd,
                                              T raw, T& v, T& l, bool& bad) {
  bool wrap_bad = false;
  v = raw;
  if (wrap) v = wrap_one_period(v, Pd, wrap_bad);
  const int i = min(max(t_floor_to_int((v - C.x_first) * C.inv_h), 0), nk - 2) + 1;
  l = ks[i - 1];
  const T h = ks[i];
  // exact check of the guess against the knots; NaN fails it
  const bool ok = (v >= l || i == 1) && (v < h || i == nk - 1);
  bad = wrap_bad || !ok;
  return i;
}

template <typename T>
__device__ __forceinline__ unsigned st_fill_bits(const ipx_axis_params& pa, const StAxC<T>& C, int i, int nk, T v) {
  unsigned f = 0;
  // a query outside the knots sits in an end cell: the compares run only there
  if (i == 1 || i == nk - 1) {
    f |= (!pa.keep_lo && v < C.x_first) ? 1u : 0u;
    f |= (!pa.keep_hi && v > C.x_last) ? 2u : 0u;
  }
  return f;
}

template <typename T, int ND>
__device__ __forceinline__ T st_apply_fill(const ipx_params& P, unsigned fill, T res) {
  // extrapolation fills, x then y then z, later axes win (_spline.py:1101-1103)
  if (fill) {
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      if (fill & (1u << (2 * ax))) res = T(P.axis[ax].fill_lo);
      if (fill & (2u << (2 * ax))) res = T(P.axis[ax].fill_hi);
    }
  }
  return res;
}

// One query, start to finish (locate, taps, 4^d contraction); returns the un-filled value.
template <typename T, int ND, bool ORD0>
__device__ __forceinline__ T st_eval_one(const StArgs<T, ND>& a, const ipx_params& P, const StShared<T, ND>& S,
                                         const T (&raw)[ND], int (&ii)[ND], unsigned& fill) {
  T tp[ND][4];
  fill = 0;
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    const int nk = a.ax[ax].nk;
    const T Pd = T(fabs(P.axis[ax].period));
    T v, l;
    bool bad;
    int i = st_locate_fast(S.ks[ax], nk, S.axc[ax], a.ax[ax].wrap != 0, Pd, raw[ax], v, l, bad);
    if (bad) {
      v = raw[ax];
      if (a.ax[ax].wrap) v = py_mod_slow(v, Pd);
      i = bin_index_slow(S.ks[ax], nk, v, S.axc[ax].x_first, S.axc[ax].inv_h);
      l = S.ks[ax][i - 1];
    }
    ii[ax] = i;
    if (!a.ax[ax].wrap) fill |= st_fill_bits(P.axis[ax], S.axc[ax], i, nk, v) << (2 * ax);
    st_taps<T, ORD0>(S.ks[ax], S.ac[ax], i, l, S.rr[ax][i], v, P.axis[ax].derivative, tp[ax]);
  }
  const T* base = a.tab;
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) base += (int64_t)(ii[ax] - 1) * a.ax[ax].stride;
  T acc = T(0);
  if constexpr (ND == 1) {
    acc = tp[0][0] * __ldg(base);
#pragma unroll
    for (int j = 1; j < 4; ++j) acc = fma(tp[0][j], __ldg(base + j), acc);
  } else if constexpr (ND == 2) {
#pragma unroll
    for (int aa = 0; aa < 4; ++aa) {
      const T* p = base + aa * a.ax[0].stride;
      T rv = tp[1][0] * __ldg(p);
#pragma unroll
      for (int j = 1; j < 4; ++j) rv = fma(tp[1][j], __ldg(p + j), rv);
      acc = fma(tp[0][aa], rv, acc);
    }
  } else {
#pragma unroll
    for (int aa = 0; aa < 4; ++aa) {
      T sy = T(0);
#pragma unroll
      for (int bb = 0; bb < 4; ++bb) {
        const T* p = base + aa * a.ax[0].stride + bb * a.ax[1].stride;
        T rv = tp[2][0] * __ldg(p);
#pragma unroll
        for (int j = 1; j < 4; ++j) rv = fma(tp[2][j], __ldg(p + j), rv);
        sy = fma(tp[1][bb], rv, sy);
      }
      acc = fma(tp[0][aa], sy, acc);
    }
  }
  return acc;
}

template <typename T, int ND>
__device__ __forceinline__ void st_views(const StArgs<T, ND>& a, const char* smem, const StAxC<T>* axc,
                                         StShared<T, ND>& S) {
  const StLayout<ND> lay = st_layout<T, ND>(a);
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    S.ks[ax] = reinterpret_cast<const T*>(smem + lay.ks_off[ax]);
    S.rr[ax] = reinterpret_cast<const T*>(smem + lay.rr_off[ax]);
    S.ac[ax] = reinterpret_cast<const Coef2<T>*>(smem + lay.ac_off[ax]);
  }
  S.axc = axc;
}

// The rare slots of the dense kernel, out of line: the queries of the quad named in `mask`
// go through the one-query path (locate with the slow paths, taps, contraction, fills) and are
// stored (overwriting what the hot path wrote for them).
template <typename T, int ND, bool ORD0>
__device__ __noinline__ void st_redo(const StArgs<T, ND>* ap, const ipx_params* Pp, const char* smem,
                                     const StAxC<T>* axc, int64_t qb, unsigned mask) {
  const StArgs<T, ND>& a = *ap;
  StShared<T, ND> S;
  st_views<T, ND>(a, smem, axc, S);
#pragma unroll 1
  for (int s = 0; s < 4; ++s) {
    const int64_t qi = qb + s;
    if (!((mask >> s) & 1u) || qi >= a.nq) continue;
    T raw[ND];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) raw[ax] = a.ax[ax].q[qi];
    int ii[ND];
    unsigned fill;
    const T res = st_eval_one<T, ND, ORD0>(a, *Pp, S, raw, ii, fill);
    a.out[qi] = st_apply_fill<T, ND>(*Pp, fill, res);
    if (a.idx_out != nullptr) {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) a.idx_out[(int64_t)ax * a.nq + qi] = ii[ax];
    }
  }
}

__device__ __forceinline__ void st_load4(const double* p, double (&r)[4]) {
  const double2 u = __ldcs(reinterpret_cast<const double2*>(p));
  const double2 w = __ldcs(reinterpret_cast<const double2*>(p) + 1);
  r[0] = u.x; r[1] = u.y; r[2] = w.x; r[3] = w.y;
}
__device__ __forceinline__ void st_load4(const float* p, float (&r)[4]) {
  const float4 u = __ldcs(reinterpret_cast<const float4*>(p));
  r[0] = u.x; r[1] = u.y; r[2] = u.z; r[3] = u.w;
}
__device__ __forceinline__ void st_store4(double* p, const double (&r)[4]) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(r[0], r[1]));
  __stcs(reinterpret_cast<double2*>(p) + 1, make_double2(r[2], r[3]));
}
__device__ __forceinline__ void st_store4(float* p, const float (&r)[4]) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(r[0], r[1], r[2], r[3]));
}

#ifndef IPX_ST_SPARSE_THREADS_F64
#define IPX_ST_SPARSE_THREADS_F64 1024  // 64 registers, no spills: C5 f64 3.91 -> 3.58 ms per 1.25e8, C4 sparse 2.86 -> 2.53 (768 threads of 80 until the end of round 2)
#endif
#ifndef IPX_ST_DENSE_THREADS_F64
#define IPX_ST_DENSE_THREADS_F64 384
#endif
template <typename T> constexpr int st_sparse_threads() { return sizeof(T) == 8 ? IPX_ST_SPARSE_THREADS_F64 : 1024; }
template <typename T> constexpr int st_dense_threads() { return sizeof(T) == 8 ? IPX_ST_DENSE_THREADS_F64 : 640; }

// Order in which the rounds (32 x Q consecutive queries each) are handed out: CHUNKS of
// nwarps * kChunkRounds consecutive rounds go to the CTAs in turn (neighbouring CTAs work on
// neighbouring stretches of the stream at the same time, so what they share is live in L2),
// and inside a chunk the warps of a CTA take the rounds in turn.  In a cell-sorted stream a
// warp's next round is then about one cell column further on, whose table rows are 3/4 the
// rows it has just used -- still in the SM's L1.
constexpr int kChunkRounds = 16;
struct StRounds {
  int64_t rounds, chunk, nchunks, ch;
  int nwarps, warp;
  __device__ __forceinline__ StRounds(int64_t r) {
    rounds = r;
    nwarps = blockDim.x >> 5;
    warp = threadIdx.x >> 5;
    chunk = (int64_t)nwarps * kChunkRounds;
    nchunks = (rounds + chunk - 1) / chunk;
    ch = blockIdx.x;
  }
  // first round of this warp, or -1
  __device__ __forceinline__ int64_t begin() {
    for (; ch < nchunks; ch += gridDim.x) {
      const int64_t r = ch * chunk + warp;
      if (r < rounds) return r;
    }
    return -1;
  }
  // the round after rnd, or -1
  __device__ __forceinline__ int64_t next(int64_t rnd) {
    rnd += nwarps;
    if (rnd < min((ch + 1) * chunk, rounds)) return rnd;
    ch += gridDim.x;
    return begin();
  }
};

__device__ __forceinline__ void st_cp_async16(unsigned smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void st_cp_async_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
// ---- bulk asynchronous copies (the TMA unit's 1-D form) completing on an mbarrier ----
__device__ __forceinline__ void st_mbar_init(unsigned mbar_s, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar_s), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(unsigned mbar_s, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar_s), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(unsigned mbar_s, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "IPX_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra IPX_MBAR_DONE;\n"
      "bra IPX_MBAR_WAIT;\n"
      "IPX_MBAR_DONE:\n"
      "}\n" ::"r"(mbar_s), "r"(phase) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void st_bulk_g2s(unsigned smem_dst, const void* gmem_src, unsigned bytes, unsigned mbar_s) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_dst),
               "l"(gmem_src), "r"(bytes), "r"(mbar_s)
               : "memory");
}
#ifndef IPX_ST_BULK
#define IPX_ST_BULK 1  // 1: table rows travel by cp.async.bulk (UBLKCP) + mbarrier; 0: by cp.async (LDGSTS)
#endif

__device__ __forceinline__ void st_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// wait until at most N of the most recently committed groups are still in flight
template <int N> __device__ __forceinline__ void st_cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------
// sparse shape: one query per lane
// DEVP: traced operands are read from device memory.  ORD0: every derivative order is 0 (known
// on the host when the operands are host-resident).
// ------------------------------------------------------------------------------------
template <typename T, int ND, bool DEVP, bool ORD0>
__global__ void __launch_bounds__(st_sparse_threads<T>(), 1)
stencil_sparse_kernel(const __grid_constant__ StArgs<T, ND> a) {
  extern __shared__ __align__(32) char smem[];
  __shared__ StAxC<T> AXC[ND];
  __shared__ ipx_params params_s;  // only used when the traced operands live in device memory
  if (a.flag != nullptr && ((a.accept >> (*a.flag & 31)) & 1u) == 0) return;  // another candidate serves this stream
  if (DEVP && threadIdx.x == 32) params_s = *a.dp;
  const ipx_params& P = DEVP ? params_s : a.hp;
  StShared<T, ND> S;
  st_stage<T, ND>(a, smem, AXC, S);

  const int lane = threadIdx.x & 31;
  const bool want_idx = a.idx_out != nullptr;
  StRounds R((a.nq + 31) / 32);
  for (int64_t rnd = R.begin(); rnd >= 0; rnd = R.next(rnd)) {
    int64_t qi = rnd * 32 + lane;
    if (qi >= a.nq) continue;
    T raw[ND];
    if (a.order != nullptr) {
      qi = (int64_t)__ldcs(a.order + qi);
      constexpr int R = ND == 3 ? 4 : ND;
      const T* r = a.recs + qi * R;
      if constexpr (R == 1) {
        raw[0] = __ldg(r);
      } else if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int m = 0; m < ND; m += 2) {
          const double2 v = __ldg(reinterpret_cast<const double2*>(r + m));
          raw[m] = T(v.x);
          if (m + 1 < ND) raw[m + 1 < ND ? m + 1 : m] = T(v.y);
        }
      } else if constexpr (R == 2) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(r));
        raw[0] = T(v.x); raw[1] = T(v.y);
      } else {
        const float4 v = __ldg(reinterpret_cast<const float4*>(r));
        raw[0] = T(v.x); raw[1] = T(v.y); raw[ND - 1] = T(v.z);
      }
    } else {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) raw[ax] = __ldcs(a.ax[ax].q + qi);
    }
    int ii[ND];
    unsigned fill;
    const T res = st_eval_one<T, ND, ORD0>(a, P, S, raw, ii, fill);
    __stcs(a.out + qi, st_apply_fill<T, ND>(P, fill, res));
    if (want_idx) {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) a.idx_out[(int64_t)ax * a.nq + qi] = ii[ax];
    }
  }
}

// ------------------------------------------------------------------------------------
// dense shape: four consecutive queries per lane share the rows of the neighbourhood.
//
// Per round a warp takes 128 consecutive queries, a quad per lane.  Lanes are grouped by cell
// column (interval indices on all axes but the last); the 4^(d-1) table rows of the group's
// column -- the stretch of the last axis its quads touch, at most kStWin values -- travel
// global -> shared with cp.async while the tap weights are formed, and every lane then reads its
// window of TAPS values per row from shared memory ONCE for its four queries.  A warp in a
// cell-sorted stream is one group (two when the round crosses into the next column).
// The hot path assumes a complete, 32-byte aligned quad whose slots are located by the first
// guess, share one cell column, sit within TAPS - 4 cells of each other along the last axis
// and in regular cells (st_stage).  Slots for which that does not hold -- and NaN results -- are
// redone by st_redo after the store, so the loop stays small.
// WM: set of axes whose queries are wrapped, as a compile-time mask (-1: read the flags).
// ------------------------------------------------------------------------------------
constexpr int kStWin = 48;  // staged values per table row and group (a multiple of 4)

template <typename T, int ND> struct StDenseCfg {
  static constexpr int Q = 4;
  static constexpr int NR = ND == 3 ? 16 : (ND == 2 ? 4 : 1);    // rows of the neighbourhood
  static constexpr int EPC = 16 / (int)sizeof(T);                 // elements per 16-byte chunk
  static constexpr int CH = kStWin / EPC;                         // chunks per staged row
  static constexpr int ROW_BYTES = NR * kStWin * (int)sizeof(T);  // staged rows of one warp
  static constexpr int WX_BYTES = ND > 1 ? 16 * 32 * (int)sizeof(T) : 0;  // first-axis taps
  static constexpr int CQ_BYTES = ND * Q * (int)sizeof(T) * 32;   // next quad's coordinates
  static constexpr int WARP_BYTES = ROW_BYTES + WX_BYTES + CQ_BYTES + 16;  // + the warp's mbarrier
};

template <typename T, int ND, int TAPS, bool DEVP, bool ORD0, int WM>
__global__ void __launch_bounds__(st_dense_threads<T>(), 1)
stencil_dense_kernel(const __grid_constant__ StArgs<T, ND> a) {
  using Cfg = StDenseCfg<T, ND>;
  constexpr int Q = Cfg::Q;
  constexpr int LA = ND - 1;  // last (run) axis
  constexpr int NR = Cfg::NR, EPC = Cfg::EPC, CH = Cfg::CH;
#define IPXS_WRAP(AXI) (WM >= 0 ? (((WM >> (AXI)) & 1) != 0) : (a.ax[AXI].wrap != 0))
  static_assert(TAPS >= 4 && TAPS <= 7, "last-axis window");
  extern __shared__ __align__(32) char smem[];
  __shared__ StAxC<T> AXC[ND];
  __shared__ ipx_params params_s;
  if (a.flag != nullptr && ((a.accept >> (*a.flag & 31)) & 1u) == 0) return;
  if (DEVP && threadIdx.x == 32) params_s = *a.dp;
  const ipx_params& P = DEVP ? params_s : a.hp;
  StShared<T, ND> S;
  st_stage<T, ND>(a, smem, AXC, S);

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  char* const wbase = smem + st_layout<T, ND>(a).total + warp * Cfg::WARP_BYTES;
  const T* const rows = reinterpret_cast<const T*>(wbase);
  const unsigned rows_s = (unsigned)__cvta_generic_to_shared(wbase);
  // the lane's taps of the first axis wait in shared memory (slot-major per tap, lanes adjacent:
  // no bank conflicts) until the contraction needs them, one tap per block of rows
  T* const wx = reinterpret_cast<T*>(wbase + Cfg::ROW_BYTES) + lane;
  bool vec_ok = (reinterpret_cast<uintptr_t>(a.out) & 31) == 0 && a.idx_out == nullptr;
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(a.ax[ax].q) & 31) == 0;
  // which open axes can fill at all (uniform)
  bool can_fill[ND];
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) can_fill[ax] = !IPXS_WRAP(ax) && (!P.axis[ax].keep_lo || !P.axis[ax].keep_hi);
  const int pitch = (int)a.pitch;  // elements per table row

  // The coordinates of the lane's NEXT quad travel global -> shared with cp.async while this one
  // is worked on (they come from DRAM: three exposed latencies per round otherwise), into the
  // lane's own 16-byte slots: chunk k of axis ax at ((ax * CPQ + k) * 32 + lane) * 16.
  constexpr int CPQ = Q * (int)sizeof(T) / 16;  // 16-byte chunks per quad and axis
  char* const cq = wbase + Cfg::ROW_BYTES + Cfg::WX_BYTES + lane * 16;
  const unsigned cq_s = (unsigned)__cvta_generic_to_shared(cq);
  // the warp's mbarrier: completion of the bulk copies of one group pass
  const unsigned mbar_s = rows_s + Cfg::ROW_BYTES + Cfg::WX_BYTES + Cfg::CQ_BYTES;
  unsigned mbar_phase = 0;
  if (IPX_ST_BULK && lane == 0) st_mbar_init(mbar_s, 1);
  __syncwarp();
  auto stage_coords = [&](int64_t rn) {
    if (rn < 0) return;
    const int64_t q0 = rn * (32 * Q) + Q * lane;
    if (!vec_ok || q0 + Q > a.nq) return;
#pragma unroll
    for (int ax = 0; ax < ND; ++ax)
#pragma unroll
      for (int k = 0; k < CPQ; ++k)
        st_cp_async16(cq_s + (ax * CPQ + k) * (32 * 16), reinterpret_cast<const char*>(a.ax[ax].q + q0) + 16 * k);
  };
  auto staged = [&](int ax, T (&c)[Q]) {
#pragma unroll
    for (int k = 0; k < CPQ; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(cq + (ax * CPQ + k) * (32 * 16));
      const T* e = reinterpret_cast<const T*>(&v);
#pragma unroll
      for (int m = 0; m < EPC; ++m) c[k * EPC + m] = e[m];
    }
  };
  StRounds R((a.nq + 32 * Q - 1) / (32 * Q));
  int64_t rnd = R.begin();
  stage_coords(rnd);
  while (rnd >= 0) {  // (warp-uniform: the warp-level primitives below are safe)
    const int64_t qb = rnd * (32 * Q) + Q * lane;
    const int64_t nxt = R.next(rnd);
    rnd = nxt;
    st_cp_async_wait();
    // lanes without a complete aligned quad take no part in the hot path
    const bool hot = vec_ok && qb + Q <= a.nq;
    T cs[ND][Q];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      if (hot) {
        staged(ax, cs[ax]);
      } else {
#pragma unroll
        for (int s = 0; s < Q; ++s) cs[ax][s] = T(0);
      }
    }
    unsigned redo = hot ? 0u : 0xfu;  // bit s: slot s must be redone
    unsigned fill = 0;                // 2 bits per (slot, axis): bit 2 * (ND * s + ax)
    T W[Q][4];                        // taps of the middle axis (3-D only)
    unsigned key = 0;                 // cell column
    int64_t colofs = 0;               // table offset of the column's first row
    // ---- column axes: slot 0 is located, slots 1..3 are checked against its cell ----
#pragma unroll
    for (int ax = 0; ax < LA; ++ax) {
      const int nk = a.ax[ax].nk;
      T (&c)[Q] = cs[ax];
      if (IPXS_WRAP(ax)) {
        const T Pd = T(fabs(P.axis[ax].period));
#pragma unroll
        for (int s = 0; s < Q; ++s) {
          bool wb;
          c[s] = wrap_one_period(c[s], Pd, wb);
          if (wb) redo |= 1u << s;
        }
      }
      const int i = min(max(t_floor_to_int((c[0] - AXC[ax].x_first) * AXC[ax].inv_h), 0), nk - 2) + 1;
      const T l = S.ks[ax][i - 1], h = S.ks[ax][i];
      const bool first = i == 1, last = i == nk - 1;
      if (!((c[0] >= l || first) && (c[0] < h || last))) redo = 0xfu;  // the guess itself is off
#pragma unroll
      for (int s = 1; s < Q; ++s)
        if (!((c[s] >= l || first) && (c[s] < h || last))) redo |= 1u << s;
      if (can_fill[ax] && (first || last)) {
        const ipx_axis_params& pa = P.axis[ax];
#pragma unroll
        for (int s = 0; s < Q; ++s) {
          unsigned f = (!pa.keep_lo && c[s] < AXC[ax].x_first) ? 1u : 0u;
          f |= (!pa.keep_hi && c[s] > AXC[ax].x_last) ? 2u : 0u;
          fill |= f << (2 * (ND * s + ax));
        }
      }
      key = key * (unsigned)(nk + 1) + (unsigned)i;
      colofs += (int64_t)(i - 1) * a.ax[ax].stride;
      const T r = S.rr[ax][i];
      if (!st_regular(r) || (!ORD0 && P.axis[ax].derivative > 1)) redo = 0xfu;
#pragma unroll
      for (int s = 0; s < Q; ++s) {
        T tp[4];
        st_taps_regular<T, ORD0>(l, r, c[s], P.axis[ax].derivative, tp);
        if (ax == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) wx[(j * 4 + s) * 32] = tp[j];
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) W[s][j] = tp[j];
        }
      }
    }
    // ---- last axis: every slot located ----
    int ii[Q];
    T lz[Q];
    {
      const int nk = a.ax[LA].nk;
      T (&c)[Q] = cs[LA];
      const T Pd = T(fabs(P.axis[LA].period));
#pragma unroll
      for (int s = 0; s < Q; ++s) {
        bool bad;
        T v;
        ii[s] = st_locate_fast(S.ks[LA], nk, AXC[LA], IPXS_WRAP(LA), Pd, c[s], v, lz[s], bad);
        c[s] = v;
        if (bad) redo |= 1u << s;
      }
      if (can_fill[LA]) {
        const ipx_axis_params& pa = P.axis[LA];
#pragma unroll
        for (int s = 0; s < Q; ++s) fill |= st_fill_bits(pa, AXC[LA], ii[s], nk, c[s]) << (2 * (ND * s + LA));
      }
    }
    const int kmin = min(min(ii[0], ii[1]), min(ii[2], ii[3]));

    // ---- groups of lanes by cell column; begin_pass issues the copy of the next group's rows ----
    if (redo == 0xfu) key = 0xffffffffu;
    unsigned todo = __ballot_sync(0xffffffffu, key != 0xffffffffu);
    unsigned fitm = 0;
    bool fit = false;
    int woff = 0;  // the lane's window starts at this element of the staged rows
    auto begin_pass = [&]() {
      const int ld = __ffs(todo) - 1;
      const unsigned kl = __shfl_sync(0xffffffffu, key, ld);
      const bool in = key == kl && ((todo >> lane) & 1u);
      const int zlo = __reduce_min_sync(0xffffffffu, in ? kmin : 0x7fffffff);
      const int zs = (zlo - 1) & ~(EPC - 1);  // 16-byte aligned start of the staged stretch
      woff = kmin - 1 - zs;
      fit = in && woff + TAPS <= kStWin;
      fitm = __ballot_sync(0xffffffffu, fit);
      const int64_t co = __shfl_sync(0xffffffffu, colofs, ld);
      // 32 / NR lanes copy one row, 16 bytes at a time, only as far as the group's quads reach
      // (and not past the row's end)
      const int zhi = __reduce_max_sync(0xffffffffu, fit ? kmin : 0);
      const int nch = min((zhi - 1 - zs + TAPS + EPC - 1) / EPC, (pitch - zs) / EPC);
#if IPX_ST_BULK
      // one bulk copy per table row (the TMA unit moves the bytes; no per-lane address loop),
      // all of them completing on the warp's mbarrier
      if (lane == 0) st_mbar_expect_tx(mbar_s, (unsigned)(NR * nch * 16));
      __syncwarp();
      if (lane < NR) {
        const int r = lane;
        int64_t ro = 0;
        if (ND == 3) ro = (int64_t)(r >> 2) * a.ax[0].stride + (int64_t)(r & 3) * a.ax[1].stride;
        if (ND == 2) ro = (int64_t)r * a.ax[0].stride;
        st_bulk_g2s(rows_s + r * (CH * 16), a.tab + co + ro + zs, (unsigned)(nch * 16), mbar_s);
      }
#else
      constexpr int LPR = 32 / NR;  // lanes per row
      const int r = lane / LPR, k0 = lane % LPR;
      int64_t ro = 0;
      if (ND == 3) ro = (int64_t)(r >> 2) * a.ax[0].stride + (int64_t)(r & 3) * a.ax[1].stride;
      if (ND == 2) ro = (int64_t)r * a.ax[0].stride;
      const T* src = a.tab + co + ro + zs + k0 * EPC;
      unsigned dst = rows_s + (r * CH + k0) * 16;
      for (int k = k0; k < nch; k += LPR) {
        st_cp_async16(dst, src);
        src += LPR * EPC;
        dst += LPR * 16;
      }
#endif
      st_cp_async_commit();
    };
    if (todo) begin_pass();
    // the next quad's coordinates follow as a second group (a lane's slots are read and written
    // by that lane only, and were read above): the contraction waits for the rows, not for these
    stage_coords(nxt);
    st_cp_async_commit();
    bool first_pass = true;

    // ---- last-axis taps, zero-padded inside the quad's window (while the copy is in flight) ----
    T WZ[Q][TAPS];
#pragma unroll
    for (int s = 0; s < Q; ++s) {
      const int o = ii[s] - kmin;
      const T r = S.rr[LA][ii[s]];
      if (o > TAPS - 4 || !st_regular(r) || (!ORD0 && P.axis[LA].derivative > 1)) redo |= 1u << s;
      T tp[4];
      st_taps_regular<T, ORD0>(lz[s], r, cs[LA][s], P.axis[LA].derivative, tp);
#pragma unroll
      for (int j = 0; j < TAPS; ++j) {
        T w = T(0);
#pragma unroll
        for (int oo = 0; oo <= TAPS - 4; ++oo)
          if (j - oo >= 0 && j - oo < 4) w = (o == oo) ? tp[j - oo] : w;
        WZ[s][j] = w;
      }
    }

    // ---- contraction: every row of the neighbourhood read once for the four slots ----
    T acc[Q];
#pragma unroll
    for (int s = 0; s < Q; ++s) acc[s] = T(0);
    while (todo) {
#if IPX_ST_BULK
      st_mbar_wait(mbar_s, mbar_phase);
      mbar_phase ^= 1u;
#else
      if (first_pass) st_cp_async_wait_group<1>(); else st_cp_async_wait();
#endif
      first_pass = false;
      __syncwarp();
      if (fit) {
        const T* win = rows + woff;
        auto row = [&](const T* p, T (&rv)[Q]) {
          T t[TAPS];
#pragma unroll
          for (int j = 0; j < TAPS; ++j) t[j] = p[j];
#pragma unroll
          for (int s = 0; s < Q; ++s) {
            T r = WZ[s][0] * t[0];
#pragma unroll
            for (int j = 1; j < TAPS; ++j) r = fma(WZ[s][j], t[j], r);
            rv[s] = r;
          }
        };
        if constexpr (ND == 1) {
          row(win, acc);
        } else if constexpr (ND == 2) {
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {
            T rv[Q];
            row(win + aa * kStWin, rv);
#pragma unroll
            for (int s = 0; s < Q; ++s) acc[s] = fma(wx[(aa * 4 + s) * 32], rv[s], acc[s]);
          }
        } else {
          // (rolled: the first-axis taps come from shared memory, so nothing here needs a static
          // index, and the loop body -- four rows -- stays within the instruction cache)
#pragma unroll 1
          for (int aa = 0; aa < 4; ++aa) {
            T sy[Q];
#pragma unroll
            for (int s = 0; s < Q; ++s) sy[s] = T(0);
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
              T rv[Q];
              row(win + (aa * 4 + bb) * kStWin, rv);
#pragma unroll
              for (int s = 0; s < Q; ++s) sy[s] = fma(W[s][bb], rv[s], sy[s]);
            }
#pragma unroll
            for (int s = 0; s < Q; ++s) acc[s] = fma(wx[(aa * 4 + s) * 32], sy[s], acc[s]);
          }
        }
      }
      __syncwarp();  // everybody is done with the staged rows
      todo &= ~fitm;
      if (todo) {
        if (fitm == 0) {
          // (cannot happen: the lane that defines the stretch always fits; never spin)
          redo |= ((todo >> lane) & 1u) ? 0xfu : 0u;
          todo = 0;
        } else {
          begin_pass();
        }
      }
    }
    if (hot) {
      // a NaN may be a non-finite table entry under a zero tap: let the one-query path decide
#pragma unroll
      for (int s = 0; s < Q; ++s)
        if (acc[s] != acc[s]) redo |= 1u << s;
      if (fill) {
#pragma unroll
        for (int s = 0; s < Q; ++s)
          acc[s] = st_apply_fill<T, ND>(P, (fill >> (2 * ND * s)) & ((1u << (2 * ND)) - 1), acc[s]);
      }
      st_store4(a.out + qb, acc);
    }
    if (redo && qb < a.nq) st_redo<T, ND, ORD0>(&a, &P, smem, AXC, qb, redo);
  }
#undef IPXS_WRAP
}

// ------------------------------------------------------------------------------------
// Stream probe.  256 quads of consecutive queries spread over the stream are located against
// the knots in global memory and the stream is classified (3/4 of the samples must agree):
//   IPX_STREAM_QUADS (3)  the four queries of a quad sit in one cell column, within the
//                         last-axis window of the dense kernel: a dense cell-sorted stream;
//   IPX_STREAM_LOCAL (1)  consecutive queries are at most one cell apart on every axis but the
//                         last: a cell-sorted stream, whatever its density;
//   IPX_STREAM_RANDOM (0) neither.
// Every candidate kernel of the call is then enqueued with the set of classes it serves and the
// ones not chosen return at once: the choice costs no host synchronisation.
// ------------------------------------------------------------------------------------
#define IPX_STREAM_RANDOM 0
#define IPX_STREAM_LOCAL 1
#define IPX_STREAM_QUADS 3
template <typename T, int ND, int TAPS>
__global__ void __launch_bounds__(256) stencil_probe_kernel(const __grid_constant__ StArgs<T, ND> a, int32_t* flag) {
  const ipx_params& P = a.dp ? *a.dp : a.hp;
  const int64_t quads = a.nq / 4;
  bool quad = false, local = false;
  if (quads > 0) {
    const int64_t qb = 4 * ((quads * (int64_t)threadIdx.x) / 256);
    int i0[ND], ip[ND];
    int kmin = 0, kmax = 0;
    quad = local = true;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) {
        const StAxis<T>& A = a.ax[ax];
        T v = A.q[qb + s];
        if (A.wrap) v = py_mod(v, T(fabs(P.axis[ax].period)));
        const T xf = A.knots[0], span = A.knots[A.nk - 1] - xf;
        T lo, hi;
        const int i = bin_index(A.knots, A.nk, v, xf, span > T(0) ? T(A.nk - 1) / span : T(0), lo, hi);
        if (s == 0) i0[ax] = i;
        if (ax < ND - 1) {
          quad = quad && i == i0[ax];
          if (s > 0) {
            // neighbours, also across the periodic seam
            int d = i - ip[ax];
            if (d < 0) d = -d;
            local = local && (d <= 1 || (A.wrap && d >= A.nk - 3));
          }
        } else {
          kmin = s == 0 ? i : min(kmin, i);
          kmax = s == 0 ? i : max(kmax, i);
        }
        ip[ax] = i;
      }
    }
    quad = quad && (kmax - kmin) <= TAPS - 4;
  }
  const int nquad = __syncthreads_count(quad);
  const int nlocal = __syncthreads_count(local);
  if (threadIdx.x == 0)
    *flag = nquad >= 192 ? IPX_STREAM_QUADS : (nlocal >= 192 ? IPX_STREAM_LOCAL : IPX_STREAM_RANDOM);
}

// kind of an axis: 0 open, 1 periodic (class form: wrapped rows all the way), 2 periodic pad-first
template <int ND> struct HaloArgs {
  const int32_t* __restrict__ perm[ND];
  int32_t n[ND];     // extents of f
  int32_t ext[ND];   // extents of the halo table
  int32_t kind[ND];
  int64_t pitch;     // elements per row in memory (last extent padded)
};

template <typename T, int ND>
__global__ void halo_kernel(const T* __restrict__ f, const HaloArgs<ND> h, int64_t total, T* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t r = e;
    HaloSrc src[ND];
    bool padding = false;
#pragma unroll
    for (int ax = ND - 1; ax >= 0; --ax) {
      const int64_t ext = ax == ND - 1 ? h.pitch : h.ext[ax];
      const int u = (int)(r % ext);
      r /= ext;
      padding = padding || u >= h.ext[ax];
      src[ax] = halo_src(padding ? 0 : u, h.n[ax], h.kind[ax], h.perm[ax]);
    }
    if (padding) {  // alignment padding at the end of a row
      out[e] = T(0);
      continue;
    }
    // tensor product of the per-axis rules; weights 1, or (2, -1) for a ghost
    T acc = T(0);
    for (int k0 = 0; k0 < src[0].cnt; ++k0) {
      const T w0 = src[0].cnt == 1 ? T(1) : (k0 == 0 ? T(2) : T(-1));
      if constexpr (ND == 1) {
        acc += w0 * f[src[0].idx[k0]];
      } else {
        for (int k1 = 0; k1 < src[1].cnt; ++k1) {
          const T w1 = w0 * (src[1].cnt == 1 ? T(1) : (k1 == 0 ? T(2) : T(-1)));
          const int64_t i01 = (int64_t)src[0].idx[k0] * h.n[1] + src[1].idx[k1];
          if constexpr (ND == 2) {
            acc += w1 * f[i01];
          } else {
            for (int k2 = 0; k2 < src[2].cnt; ++k2) {
              const T w2 = w1 * (src[2].cnt == 1 ? T(1) : (k2 == 0 ? T(2) : T(-1)));
              acc += w2 * f[i01 * h.n[2] + src[2].idx[k2]];
            }
          }
        }
      }
    }
    out[e] = acc;
  }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct StDevice {
  int dev;
  int sms;
  int smem_optin;
};

// per-device constants, looked up once per device (immutable afterwards)
static bool st_device(StDevice& d) {
  constexpr int kMaxDev = 64;
  static std::atomic<int> ready[kMaxDev];
  static int sms[kMaxDev], optin[kMaxDev];
  if (cudaGetDevice(&d.dev) != cudaSuccess) return false;
  if (d.dev < 0 || d.dev >= kMaxDev) {
    return cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, d.dev) == cudaSuccess &&
           cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.dev) == cudaSuccess;
  }
  if (ready[d.dev].load(std::memory_order_acquire) == 0) {
    int s_ = 0, o_ = 0;
    if (cudaDeviceGetAttribute(&s_, cudaDevAttrMultiProcessorCount, d.dev) != cudaSuccess) return false;
    if (cudaDeviceGetAttribute(&o_, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.dev) != cudaSuccess) return false;
    sms[d.dev] = s_;
    optin[d.dev] = o_;
    ready[d.dev].store(1, std::memory_order_release);
  }
  d.sms = sms[d.dev];
  d.smem_optin = optin[d.dev];
  return true;
}

// The opt-in shared-memory size of a kernel is a per-(kernel, device) attribute: raise it when a
// launch needs more than any earlier one (never lowered), not on every launch.
static bool st_grant_smem(const void* kern, int dev, size_t smem) {
  static std::mutex mu;
  static std::unordered_map<unsigned long long, size_t> granted;
  const unsigned long long key = (unsigned long long)(uintptr_t)kern * 64ull + (unsigned long long)(dev & 63);
  std::lock_guard<std::mutex> lock(mu);
  auto it = granted.find(key);
  if (it != granted.end() && it->second >= smem) return true;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
  granted[key] = smem;
  return true;
}

}  // namespace ipx
#include "ipx_stencil_col.cuh"
namespace ipx {

template <typename T, int ND, int TAPS, int WM>
static void (*st_dense_entry(const StArgs<T, ND>& a, bool ord0))(const StArgs<T, ND>) {
  return a.dp ? stencil_dense_kernel<T, ND, TAPS, true, false, WM>
              : (ord0 ? stencil_dense_kernel<T, ND, TAPS, false, true, WM>
                      : stencil_dense_kernel<T, ND, TAPS, false, false, WM>);
}

template <typename T, int ND, bool DENSE, int TAPS>
static int st_launch_q(const StArgs<T, ND>& a, const StDevice& d, cudaStream_t s) {
  const StLayout<ND> lay = st_layout<T, ND>(a);
  int threads = DENSE ? st_dense_threads<T>() : st_sparse_threads<T>();
  size_t smem = (size_t)lay.total;
  if constexpr (DENSE) {
    // per warp: staged rows, first-axis taps, the next quad's coordinates; fewer warps if the
    // knot vectors leave less room
    const long long room = (long long)d.smem_optin - 2048 - (long long)lay.total;
    long long warps = room / StDenseCfg<T, ND>::WARP_BYTES;
    if (warps > threads / 32) warps = threads / 32;
    if (warps < 4) return IPX_EUNSUPPORTED;
    threads = (int)warps * 32;
    smem += (size_t)warps * StDenseCfg<T, ND>::WARP_BYTES;
  } else if ((long long)smem + 2048 > (long long)d.smem_optin) {
    return IPX_EUNSUPPORTED;
  }
  bool ord0 = a.dp == nullptr;
  for (int ax = 0; ax < ND && ord0; ++ax) ord0 = a.hp.axis[ax].derivative <= 0;
  void (*kern)(const StArgs<T, ND>);
  if constexpr (DENSE) {
    // the wrapped axes as a compile-time mask where it is one of the common ones: none, or all
    // but the last axis
    int mask = 0;
    for (int ax = 0; ax < ND; ++ax) mask |= (a.ax[ax].wrap != 0) << ax;
    constexpr int HEAD = (1 << (ND - 1)) - 1;
    if (mask == 0) kern = st_dense_entry<T, ND, TAPS, 0>(a, ord0);
    else if (ND > 1 && mask == HEAD) kern = st_dense_entry<T, ND, TAPS, HEAD>(a, ord0);
    else kern = st_dense_entry<T, ND, TAPS, -1>(a, ord0);
  } else {
    kern = a.dp ? stencil_sparse_kernel<T, ND, true, false>
                : (ord0 ? stencil_sparse_kernel<T, ND, false, true> : stencil_sparse_kernel<T, ND, false, false>);
  }
  // (static shared memory counts towards the 48 KB default limit too)
  if (smem > 32 * 1024 && !st_grant_smem(reinterpret_cast<const void*>(kern), d.dev, smem)) return IPX_ELAUNCH;
  const int qpt = DENSE ? 4 : 1;
  int64_t need = (a.nq + (int64_t)threads * qpt - 1) / ((int64_t)threads * qpt);
  int64_t grid = d.sms;  // persistent: one CTA per SM
  if (need < grid) grid = need;
  kern<<<(int)grid, threads, smem, s>>>(a);
  count_launches(1);
  const cudaError_t err = cudaPeekAtLastError();
  if (err != cudaSuccess && getenv("IPX_DEBUG") != nullptr)
    fprintf(stderr, "ipx stencil launch (dense %d, %d threads, %zu B shared): %s\n", (int)DENSE, threads, smem,
            cudaGetErrorString(err));
  return err == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

#ifndef IPX_STENCIL_TAPS
#define IPX_STENCIL_TAPS 6
#endif

// the record form of the same call (ipx_eval*d_*: tile kernel / generic gathers on AoS records)
static int eval_records(const float* const* q, int64_t nq, const float* const* k, const int32_t* n, int nd,
                        const float* rec, int32_t mask, const ipx_params* p, int32_t pod, float* out, void* st) {
  const float* tabs[1] = {rec};
  if (nd == 1) return ipx_eval1d_f32(q[0], nq, k[0], nullptr, n[0], tabs, IPX_AOS, 1, IPX_CUBIC, mask, p, pod, out, nullptr, st);
  if (nd == 2)
    return ipx_eval2d_f32(q[0], q[1], nq, k[0], k[1], nullptr, nullptr, n[0], n[1], tabs, IPX_AOS, 1, IPX_CUBIC, mask,
                          p, pod, out, nullptr, st);
  return ipx_eval3d_f32(q[0], q[1], q[2], nq, k[0], k[1], k[2], nullptr, nullptr, nullptr, n[0], n[1], n[2], tabs,
                        IPX_AOS, 1, IPX_CUBIC, mask, p, pod, out, nullptr, st);
}
static int eval_records(const double* const* q, int64_t nq, const double* const* k, const int32_t* n, int nd,
                        const double* rec, int32_t mask, const ipx_params* p, int32_t pod, double* out, void* st) {
  const double* tabs[1] = {rec};
  if (nd == 1) return ipx_eval1d_f64(q[0], nq, k[0], nullptr, n[0], tabs, IPX_AOS, 1, IPX_CUBIC, mask, p, pod, out, nullptr, st);
  if (nd == 2)
    return ipx_eval2d_f64(q[0], q[1], nq, k[0], k[1], nullptr, nullptr, n[0], n[1], tabs, IPX_AOS, 1, IPX_CUBIC, mask,
                          p, pod, out, nullptr, st);
  return ipx_eval3d_f64(q[0], q[1], q[2], nq, k[0], k[1], k[2], nullptr, nullptr, nullptr, n[0], n[1], n[2], tabs,
                        IPX_AOS, 1, IPX_CUBIC, mask, p, pod, out, nullptr, st);
}

// tables up to about the size of L2 (126 MB): random-order gathers of the 4^d neighbourhood mostly
// hit L2; beyond that the record form touches fewer DRAM sectors per query (2^d aligned records)
constexpr double kL2TableBytes = 160e6;

template <typename T, int ND>
static int launch_stencil(const T* const* q, int64_t nq, const T* const* knots, const T* const* slope_knots,
                          const int32_t* n, const T* halo, const T* records, int32_t slope_method, double c,
                          int32_t periodic_mask, const ipx_params* params, int32_t params_on_device, T* out,
                          int32_t* idx_out, int32_t hint, int32_t* workspace, void* stream,
                          const uint32_t* order = nullptr, const T* coord_recs = nullptr) {
  if (nq < 0 || q == nullptr || knots == nullptr || n == nullptr || params == nullptr) return IPX_EINVAL;
  if ((order == nullptr) != (coord_recs == nullptr)) return IPX_EINVAL;
  if (slope_method != IPX_SLOPE_CUBIC && slope_method != IPX_SLOPE_CARDINAL) return IPX_EUNSUPPORTED;
  if (hint < IPX_STENCIL_AUTO || hint > IPX_STENCIL_COLUMN) return IPX_EINVAL;
  if (nq == 0) return IPX_OK;
  if (out == nullptr || halo == nullptr) return IPX_EINVAL;
  StArgs<T, ND> a;
  int64_t stride = 1;
  double cells = 1.0, halo_bytes = (double)sizeof(T);
  int32_t rec_mask = 0, rec_n[ND];
  for (int ax = ND - 1; ax >= 0; --ax) {
    if (q[ax] == nullptr || knots[ax] == nullptr || n[ax] < 2) return IPX_EINVAL;
    const int per = (periodic_mask >> ax) & 1;             // class form: slopes on the un-padded knots
    const int padfirst = (periodic_mask >> (4 + ax)) & 1;  // functional form: slopes on the padded knots
    if (per && padfirst) return IPX_EINVAL;
    StAxis<T>& A = a.ax[ax];
    A.q = q[ax];
    A.knots = knots[ax];
    A.nk = (per || padfirst) ? n[ax] + 2 : n[ax];
    A.wrap = per | padfirst;
    A.ring = per;
    if (per) {
      if (slope_knots == nullptr || slope_knots[ax] == nullptr) return IPX_EINVAL;
      A.sknots = slope_knots[ax];
      A.ns = n[ax];
    } else {
      A.sknots = knots[ax];
      A.ns = A.nk;
    }
    A.stride = stride;
    // (the last axis of the halo table is padded to a multiple of 32 bytes: rows start aligned)
    stride *= ax == ND - 1 ? st_halo_pitch<T>(A.nk + 2) : (int64_t)A.nk + 2;
    cells *= (double)(A.nk - 1);
    halo_bytes *= (double)(A.nk + 2);
    // the record form of the same axis: un-padded records + knot mapping (class form), or
    // records of the padded table (functional form)
    rec_mask |= per ? IPX_PERIODIC(ax) : (padfirst ? IPX_WRAP_ONLY(ax) : 0);
    rec_n[ax] = padfirst ? n[ax] + 2 : n[ax];
  }
  if ((reinterpret_cast<uintptr_t>(halo) & 15) != 0) return IPX_EINVAL;
  a.tab = halo;
  a.pitch = st_halo_pitch<T>((int64_t)a.ax[ND - 1].nk + 2);
  a.out = out;
  a.idx_out = idx_out;
  a.nq = nq;
  a.cardinal = slope_method == IPX_SLOPE_CARDINAL;
  a.fac = T(1) - T(c);
  a.flag = nullptr;
  a.accept = 0;
  a.order = order;
  a.recs = coord_recs;
  if (params_on_device) {
    a.dp = params;
    a.hp = ipx_params{};
  } else {
    a.dp = nullptr;
    a.hp = *params;
  }
  StDevice d;
  if (!st_device(d)) return IPX_ELAUNCH;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (order != nullptr) {
    if ((reinterpret_cast<uintptr_t>(coord_recs) & 31) != 0) return IPX_EINVAL;
    return st_launch_q<T, ND, false, 4>(a, d, s);  // the shape that takes its queries one per lane
  }
  bool low_order = true;  // (the dense shape's hot loop covers derivative orders 0 and 1)
  if (!params_on_device)
    for (int ax = 0; ax < ND; ++ax) low_order = low_order && params->axis[ax].derivative <= 1;
  // Dense streams (several queries per cell, cell-sorted) share rows between the four queries of
  // a lane; below about three queries per cell consecutive queries are rarely close enough.
  // (The dense shape does not write interval indices: such calls run the sparse shape.)
  const bool dense_ok = idx_out == nullptr && low_order;
  if (hint == IPX_STENCIL_DENSE && dense_ok) return st_launch_q<T, ND, true, IPX_STENCIL_TAPS>(a, d, s);
  if constexpr (ND == 3) {
    if (hint == IPX_STENCIL_COLUMN && dense_ok) {
      const int rc = st_launch_column<T>(a, d, s);
      if (rc != IPX_EUNSUPPORTED) return rc;
    }
  }
  if (hint != IPX_STENCIL_AUTO || workspace == nullptr || idx_out != nullptr || nq < (1 << 16))
    return st_launch_q<T, ND, false, 4>(a, d, s);

  // ---- automatic choice, on the device.  Measured on B200 (profiles/stencil_r02.jsonl):
  //   dense cell-sorted stream   3-D: tile kernel on the records (if the caller has them), else the
  //                              column shape (ipx_stencil_col.cuh); 1-D / 2-D: the dense shape
  //   cell-sorted, sparse        the sparse shape (rows shared through L1)
  //   random order               table about L2-sized or smaller, 3-D: the sparse shape (8x fewer
  //                              bytes than the records); larger tables, and 2-D (same bytes,
  //                              aligned records): the records if the caller has them
  // (in 3-D the dense shape does not beat the sparse one yet -- 2.9 vs 2.8 ms per 1e8 queries on
  // 256^3 f64 -- so the automatic choice uses it in 1-D / 2-D only; the hint still selects it)
  const bool dense = dense_ok && ND < 3 && (double)nq >= 3.0 * cells;
  const bool rec_random = records != nullptr && (ND == 2 || (ND == 3 && halo_bytes > kL2TableBytes));
  // (the tile kernel on the records overtakes the sparse shape between three and six queries per
  // cell: 256^3 f64, 5e7 queries 1.73 vs 1.44 ms, 1e8 queries 1.86 vs 2.83 ms)
  const bool rec_quads = records != nullptr && ND == 3 && dense_ok && (double)nq >= 4.0 * cells;
  // (without records, dense cell-sorted 3-D streams go to the column shape: 256^3 f64, 1e8 queries
  // 2.07 vs 2.79 ms; at 5e7 queries, three per cell, the sparse shape is still ahead, 1.46 vs 1.76)
  bool col_quads = false;
  if constexpr (ND == 3) {
    ColPlan pl;
    col_quads = !rec_quads && dense_ok && (double)nq >= 4.5 * cells && st_column_plan<T>(a, d, pl);
  }
  if (!dense && !rec_random && !rec_quads && !col_quads) return st_launch_q<T, ND, false, 4>(a, d, s);  // one candidate: no probe
  stencil_probe_kernel<T, ND, IPX_STENCIL_TAPS><<<1, 256, 0, s>>>(a, workspace);
  count_launches(1);
  a.flag = workspace;
  uint32_t sparse_accept = 1u << IPX_STREAM_LOCAL;
  if (!dense && !rec_quads && !col_quads) sparse_accept |= 1u << IPX_STREAM_QUADS;
  if (!rec_random) sparse_accept |= 1u << IPX_STREAM_RANDOM;
  uint32_t rec_accept = (rec_random ? 1u << IPX_STREAM_RANDOM : 0u) | (rec_quads ? 1u << IPX_STREAM_QUADS : 0u);
  int rc = IPX_OK;
  if (dense && !rec_quads) {
    a.accept = 1u << IPX_STREAM_QUADS;
    rc = st_launch_q<T, ND, true, IPX_STENCIL_TAPS>(a, d, s);
    if (rc != IPX_OK) return rc;
  }
  if constexpr (ND == 3) {
    if (col_quads) {
      a.accept = 1u << IPX_STREAM_QUADS;
      rc = st_launch_column<T>(a, d, s);
      if (rc != IPX_OK) return rc;
    }
  }
  a.accept = sparse_accept;
  rc = st_launch_q<T, ND, false, 4>(a, d, s);
  if (rc != IPX_OK) return rc;
  if (rec_accept) {
    g_run_if = RunIf{workspace, rec_accept};
    rc = eval_records(q, nq, knots, rec_n, ND, records, rec_mask, params, params_on_device, out, stream);
    g_run_if = RunIf{nullptr, 0u};
  }
  return rc;
}

template <typename T, int ND>
static int launch_halo_nd(const T* f, const int32_t* n, int32_t periodic_mask, const int32_t* const* perm, T* out,
                          cudaStream_t s) {
  HaloArgs<ND> h;
  int64_t total = 1;
  for (int ax = 0; ax < ND; ++ax) {
    if (n[ax] < 2) return IPX_EINVAL;
    h.n[ax] = n[ax];
    const int ring = (periodic_mask >> ax) & 1, padfirst = (periodic_mask >> (4 + ax)) & 1;
    if (ring && padfirst) return IPX_EINVAL;
    h.kind[ax] = ring ? 1 : (padfirst ? 2 : 0);
    h.ext[ax] = n[ax] + (h.kind[ax] ? 4 : 2);
    h.perm[ax] = (h.kind[ax] && perm != nullptr) ? perm[ax] : nullptr;
    total *= ax == ND - 1 ? st_halo_pitch<T>(h.ext[ax]) : h.ext[ax];
  }
  h.pitch = st_halo_pitch<T>(h.ext[ND - 1]);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (total + 255) / 256;
  int64_t cap = (int64_t)sms * 16;
  halo_kernel<T, ND><<<(int)(need < cap ? need : cap), 256, 0, s>>>(f, h, total, out);
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_halo(int32_t ndim, const T* f, const int32_t* n, int32_t periodic_mask,
                       const int32_t* const* perm, T* out, void* stream) {
  if (ndim < 1 || ndim > 3 || f == nullptr || n == nullptr || out == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ndim == 1) return launch_halo_nd<T, 1>(f, n, periodic_mask, perm, out, s);
  if (ndim == 2) return launch_halo_nd<T, 2>(f, n, periodic_mask, perm, out, s);
  return launch_halo_nd<T, 3>(f, n, periodic_mask, perm, out, s);
}

template <typename T>
static int launch_stencil_any(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,
                              const T* const* slope_knots, const int32_t* n, const T* halo, const T* records,
                              int32_t slope_method,
                              double c, int32_t periodic_mask, const ipx_params* params, int32_t params_on_device,
                              T* out, int32_t* idx_out, int32_t hint, int32_t* workspace, void* stream,
                              const uint32_t* order = nullptr, const T* coord_recs = nullptr) {
  if (ndim == 1)
    return launch_stencil<T, 1>(q, nq, knots, slope_knots, n, halo, records, slope_method, c, periodic_mask,
                                params, params_on_device, out, idx_out, hint, workspace, stream, order, coord_recs);
  if (ndim == 2)
    return launch_stencil<T, 2>(q, nq, knots, slope_knots, n, halo, records, slope_method, c, periodic_mask,
                                params, params_on_device, out, idx_out, hint, workspace, stream, order, coord_recs);
  if (ndim == 3)
    return launch_stencil<T, 3>(q, nq, knots, slope_knots, n, halo, records, slope_method, c, periodic_mask,
                                params, params_on_device, out, idx_out, hint, workspace, stream, order, coord_recs);
  return IPX_EINVAL;
}

}  // namespace ipx

extern "C" int64_t ipx_halo_table_elems(int32_t ndim, const int32_t* n, int32_t periodic_mask, int32_t elem_size) {
  if (ndim < 1 || ndim > 3 || n == nullptr || (elem_size != 4 && elem_size != 8)) return -1;
  int64_t total = 1;
  for (int ax = 0; ax < ndim; ++ax) {
    const int per = ((periodic_mask >> ax) & 1) | ((periodic_mask >> (4 + ax)) & 1);
    const int64_t ext = (int64_t)n[ax] + (per ? 4 : 2);
    total *= ax == ndim - 1 ? (elem_size == 8 ? ipx::st_halo_pitch<double>(ext) : ipx::st_halo_pitch<float>(ext)) : ext;
  }
  return total;
}
extern "C" int ipx_halo_table_f32(int32_t ndim, const float* f, const int32_t* n, int32_t periodic_mask,
                                  const int32_t* const* perm, float* out, void* stream) {
  return ipx::launch_halo<float>(ndim, f, n, periodic_mask, perm, out, stream);
}
extern "C" int ipx_halo_table_f64(int32_t ndim, const double* f, const int32_t* n, int32_t periodic_mask,
                                  const int32_t* const* perm, double* out, void* stream) {
  return ipx::launch_halo<double>(ndim, f, n, periodic_mask, perm, out, stream);
}
extern "C" int ipx_eval_stencil_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                                    const float* const* slope_knots, const int32_t* n, const float* halo,
                                    const float* records, int32_t slope_method, double c, int32_t periodic_mask,
                                    const ipx_params* params, int32_t params_on_device, float* out,
                                    int32_t* idx_out, int32_t hint, int32_t* workspace, void* stream) {
  return ipx::launch_stencil_any<float>(ndim, q, nq, knots, slope_knots, n, halo, records, slope_method, c, periodic_mask,
                                        params, params_on_device, out, idx_out, hint, workspace, stream);
}
extern "C" int ipx_eval_stencil_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                                    const double* const* slope_knots, const int32_t* n, const double* halo,
                                    const double* records, int32_t slope_method, double c, int32_t periodic_mask,
                                    const ipx_params* params, int32_t params_on_device, double* out,
                                    int32_t* idx_out, int32_t hint, int32_t* workspace, void* stream) {
  return ipx::launch_stencil_any<double>(ndim, q, nq, knots, slope_knots, n, halo, records, slope_method, c, periodic_mask,
                                         params, params_on_device, out, idx_out, hint, workspace, stream);
}

// the pre-sorted call with its permutation passes folded in (see StArgs::order)
#define IPX_DEFINE_STENCIL_PERM(SUF, T)                                                                            \
  extern "C" int ipx_eval_stencil_perm_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,   \
                                             const T* const* slope_knots, const int32_t* n, const T* halo,         \
                                             int32_t slope_method, double c, int32_t periodic_mask,                \
                                             const ipx_params* params, int32_t params_on_device,                   \
                                             const uint32_t* order, const T* coord_records, T* out,                \
                                             void* stream) {                                                       \
    if (order == nullptr || coord_records == nullptr) return IPX_EINVAL;                                           \
    return ipx::launch_stencil_any<T>(ndim, q, nq, knots, slope_knots, n, halo, nullptr, slope_method, c,          \
                                      periodic_mask, params, params_on_device, out, nullptr, IPX_STENCIL_SPARSE,   \
                                      nullptr, stream, order, coord_records);                                      \
  }
IPX_DEFINE_STENCIL_PERM(f32, float)
IPX_DEFINE_STENCIL_PERM(f64, double)
