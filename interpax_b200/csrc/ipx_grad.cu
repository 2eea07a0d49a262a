// ipx_grad.cu -- derivatives of the cubic evaluation with respect to the KNOTS and to f
// (what the reference's TestAD differentiates, tests/test_interpolate.py:542-637:
// jacfwd / jacrev of interpNd and InterpolatorND with respect to the knot vector).
//
// For the three-point stencils (method = cubic, cardinal, catmull-rom) the whole map
//     (knots, f) -> interpNd(q, knots, f)
// is local: a query sees the 4 knots i-2 .. i+1 around its cell on every axis and the 4^d
// neighbourhood of f (ipx_stencil.cu), and the value is multilinear in the per-axis tap weights,
//     out = sum_{a,b,c} Tx[a] Ty[b] Tz[c] V[a,b,c].
// So each thread evaluates the tap weights of its query in DUAL numbers seeded on the 4 local
// knots of the axis (the arithmetic of hermite_weights / st_node_coefs, reciprocal guards
// included), contracts the other axes in plain numbers, and gets
//     d out / d x_k = sum_a dTx[a]/dx_k * P_x[a],     P_x[a] = sum_{b,c} Ty[b] Tz[c] V[a,b,c]
//     d out / d V[a,b,c] = Tx[a] Ty[b] Tz[c]
// from which the two modes follow:
//   VJP (reverse)  grad_knots[ax][k] += cot[q] * d out/d x_k,  grad_f[...] += cot[q] * d out/d V
//                  (atomics; ghost rows of the halo are linear in f and fold back into f);
//   JVP (forward)  tangent[q] = sum_k xdot[ax][k] d out/d x_k + sum fdot * d out/d V.
// Queries replaced by an extrapolation fill have zero derivative, as in the reference (where the
// fill is a constant).  cubic2 (a global tridiagonal solve), monotonic and akima (non-linear,
// non-local) are not covered: IPX_EUNSUPPORTED.
#include <cuda_runtime.h>

#include "ipx_device.cuh"
#include "ipx_halo.cuh"
#include "ipx_internal.cuh"

namespace ipx {

constexpr int kGradThreads = 128;

// value + derivatives with respect to the 4 local knots of one axis
template <typename T> struct Dual4 {
  T v, d[4];
};
template <typename T> __device__ __forceinline__ Dual4<T> dconst(T v) { return Dual4<T>{v, {T(0), T(0), T(0), T(0)}}; }
template <typename T> __device__ __forceinline__ Dual4<T> dseed(T v, int k) {
  Dual4<T> r = dconst(v);
  r.d[k] = T(1);
  return r;
}
template <typename T> __device__ __forceinline__ Dual4<T> operator+(const Dual4<T>& a, const Dual4<T>& b) {
  Dual4<T> r;
  r.v = a.v + b.v;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] + b.d[k];
  return r;
}
template <typename T> __device__ __forceinline__ Dual4<T> operator-(const Dual4<T>& a, const Dual4<T>& b) {
  Dual4<T> r;
  r.v = a.v - b.v;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] - b.d[k];
  return r;
}
template <typename T> __device__ __forceinline__ Dual4<T> operator-(const Dual4<T>& a) {
  Dual4<T> r;
  r.v = -a.v;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.d[k] = -a.d[k];
  return r;
}
template <typename T> __device__ __forceinline__ Dual4<T> operator*(const Dual4<T>& a, const Dual4<T>& b) {
  Dual4<T> r;
  r.v = a.v * b.v;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.d[k] = a.d[k] * b.v + a.v * b.d[k];
  return r;
}
template <typename T> __device__ __forceinline__ Dual4<T> operator*(T s, const Dual4<T>& a) {
  Dual4<T> r;
  r.v = s * a.v;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.d[k] = s * a.d[k];
  return r;
}
// where(dx == 0, 0, 1 / dx): a constant (zero derivative) on the guarded branch
template <typename T> __device__ __forceinline__ Dual4<T> drecip(const Dual4<T>& a) {
  if (a.v == T(0)) return dconst(T(0));
  Dual4<T> r;
  r.v = T(1) / a.v;
  const T m = -r.v * r.v;
#pragma unroll
  for (int k = 0; k < 4; ++k) r.d[k] = m * a.d[k];
  return r;
}

template <typename T, int ND> struct GradArgs {
  const T* __restrict__ q[ND];
  const T* __restrict__ knots[ND];     // nk lookup knots (padded + sorted on a periodic axis)
  const T* __restrict__ sknots[ND];    // class-form periodic axis: the n original knots
  const int32_t* __restrict__ perm[ND];  // periodic: sorted position -> table row, or NULL
  int32_t n[ND], nk[ND], kind[ND];     // table extent; lookup knots; 0 open, 1 periodic (class), 2 pad-first
  const T* __restrict__ f;
  const T* __restrict__ cot;            // VJP: cotangent (nq)
  T* __restrict__ grad_knots[ND];       // VJP: n[ax] entries each, or NULL
  T* __restrict__ grad_f;               // VJP: shape of f, or NULL
  const T* __restrict__ knots_dot[ND];  // JVP: n[ax] entries each, or NULL
  const T* __restrict__ f_dot;          // JVP: shape of f, or NULL
  T* __restrict__ tangent;              // JVP: nq
  int64_t nq;
  T fac;
  int32_t cardinal, jvp;
  ipx_params hp;
  const ipx_params* dp;
};

__device__ __forceinline__ void grad_atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void grad_atomic_add(double* p, double v) { atomicAdd(p, v); }

// stencil coefficients (a, b, c) of one knot in dual numbers; lo / hi: reciprocal widths of the
// cells below / above it, wide: reciprocal of the two-cell span (cardinal)
template <typename T>
__device__ __forceinline__ void dual_coefs(bool first, bool last, bool cardinal, T fac, const Dual4<T>& lo,
                                           const Dual4<T>& hi, const Dual4<T>& wide, Dual4<T>& ca, Dual4<T>& cb,
                                           Dual4<T>& cc) {
  if (first) {
    const Dual4<T> s = cardinal ? fac * hi : hi;
    ca = dconst(T(0)); cb = -s; cc = s;
  } else if (last) {
    const Dual4<T> s = cardinal ? fac * lo : lo;
    ca = -s; cb = s; cc = dconst(T(0));
  } else if (cardinal) {
    const Dual4<T> s = fac * wide;
    ca = -s; cb = dconst(T(0)); cc = s;
  } else {
    ca = T(-0.5) * lo; cb = T(0.5) * (lo - hi); cc = T(0.5) * hi;
  }
}

template <typename T, int ND>
__global__ void __launch_bounds__(kGradThreads) grad_stencil_kernel(const GradArgs<T, ND> a) {
  __shared__ ipx_params params_s;
  if (threadIdx.x == 0) params_s = a.dp ? *a.dp : a.hp;
  __syncthreads();
  const ipx_params& P = params_s;
  constexpr int NV = ND == 3 ? 64 : (ND == 2 ? 16 : 4);
  const int64_t stride = (int64_t)gridDim.x * kGradThreads;
  for (int64_t qi = (int64_t)blockIdx.x * kGradThreads + threadIdx.x; qi < a.nq; qi += stride) {
    Dual4<T> tap[ND][4];
    int cell[ND];
    bool filled = false;
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      const T* ks = a.knots[ax];
      const int nk = a.nk[ax];
      T v = a.q[ax][qi];
      if (a.kind[ax] != 0) v = py_mod(v, T(fabs(P.axis[ax].period)));
      const T xf = ks[0], xl = ks[nk - 1], span = xl - xf;
      T lo, hi;
      const int i = bin_index(ks, nk, v, xf, span > T(0) ? T(nk - 1) / span : T(0), lo, hi);
      cell[ax] = i;
      if (a.kind[ax] == 0) {
        if (!P.axis[ax].keep_lo && v < xf) filled = true;
        if (!P.axis[ax].keep_hi && v > xl) filled = true;
      }
      // The four local knots i-2 .. i+1 (lookup positions) carry the seeds 0 .. 3.  Only their
      // DIFFERENCES and the lower knot of the cell enter: D[j] = X[j+1] - X[j].  Open and
      // pad-first axes difference the lookup knots (a difference that reaches outside the vector
      // belongs to a one-sided end and is not used).  The class form of a periodic axis takes its
      // slopes on the ORIGINAL knots, one-sided at table rows 0 and n-1, then wraps: the local
      // knots are the original ones at (i-3 .. i) mod n, differenced across the seam with the period
      // added -- the one-sided seam stencils never use that difference.
      int mL = i - 1, mR = i, ns = nk;
      bool have0 = i - 2 >= 0, have3 = i + 1 < nk;
      Dual4<T> D[3];
      if (a.kind[ax] == 1) {
        ns = a.n[ax];
        const T* xs = a.sknots[ax];
        const T Pd = T(fabs(P.axis[ax].period));
        int o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          o[j] = (i - 3 + j) % ns;
          if (o[j] < 0) o[j] += ns;
        }
        mL = o[1];
        mR = o[2];
        have0 = have3 = true;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const T dv = xs[o[j + 1]] - xs[o[j]] + (o[j + 1] <= o[j] ? Pd : T(0));
          D[j] = dseed(dv, j + 1) - dseed(T(0), j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int p0 = min(max(i - 2 + j, 0), nk - 1), p1 = min(max(i - 1 + j, 0), nk - 1);
          D[j] = dseed(ks[p1], j + 1) - dseed(ks[p0], j);
        }
      }
      const Dual4<T> X1 = dseed(ks[i - 1], 1);
      const Dual4<T> rL = have0 ? drecip(D[0]) : dconst(T(0));
      const Dual4<T> rC = drecip(D[1]);
      const Dual4<T> rR = have3 ? drecip(D[2]) : dconst(T(0));
      const Dual4<T> wL = have0 ? drecip(D[0] + D[1]) : dconst(T(0));
      const Dual4<T> wR = have3 ? drecip(D[1] + D[2]) : dconst(T(0));
      Dual4<T> aL, bL, cL, aR, bR, cR;
      dual_coefs<T>(mL == 0, mL == ns - 1, a.cardinal != 0, a.fac, rL, rC, wL, aL, bL, cL);
      dual_coefs<T>(mR == 0, mR == ns - 1, a.cardinal != 0, a.fac, rC, rR, wR, aR, bR, cR);
      // Hermite weights of _get_t_der . A_CUBIC (hermite_weights), in dual numbers
      const Dual4<T> dx = D[1];
      const Dual4<T> t = (dconst(v) - X1) * rC;
      const int order = P.axis[ax].derivative;
      Dual4<T> tt0, tt1, tt2, tt3;
      if (order <= 0) {
        tt0 = dconst(T(1)); tt1 = t; tt2 = t * t; tt3 = tt2 * t;
      } else if (order == 1) {
        tt0 = dconst(T(0)); tt1 = rC; tt2 = (T(2) * t) * rC; tt3 = (T(3) * (t * t)) * rC;
      } else if (order == 2) {
        const Dual4<T> s = rC * rC;
        tt0 = dconst(T(0)); tt1 = dconst(T(0)); tt2 = T(2) * s; tt3 = (T(6) * t) * s;
      } else if (order == 3) {
        tt0 = dconst(T(0)); tt1 = dconst(T(0)); tt2 = dconst(T(0)); tt3 = T(6) * (rC * rC * rC);
      } else {
        tt0 = tt1 = tt2 = tt3 = dconst(T(0));
      }
      const Dual4<T> w0 = tt0 - T(3) * tt2 + T(2) * tt3;
      const Dual4<T> w1 = T(3) * tt2 - T(2) * tt3;
      const Dual4<T> w2 = (tt1 - T(2) * tt2 + tt3) * dx;
      const Dual4<T> w3 = (tt3 - tt2) * dx;
      tap[ax][0] = w2 * aL;
      tap[ax][1] = w0 + w2 * bL + w3 * aR;
      tap[ax][2] = w1 + w2 * cL + w3 * bR;
      tap[ax][3] = w3 * cR;
    }
    if (filled) {
      if (a.jvp) a.tangent[qi] = T(0);
      continue;
    }
    // sources of the 4 rows per axis (ghost rows fold into two table rows)
    HaloSrc src[ND][4];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax)
#pragma unroll
      for (int j = 0; j < 4; ++j) src[ax][j] = halo_src(cell[ax] - 1 + j, a.n[ax], a.kind[ax], a.perm[ax]);
    // partial contractions P[ax][j] = sum over the other axes of (their taps) * V
    T part[ND][4];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax)
#pragma unroll
      for (int j = 0; j < 4; ++j) part[ax][j] = T(0);
    const T cot = a.jvp ? T(0) : a.cot[qi];
    T tan = T(0);
    for (int e = 0; e < NV; ++e) {
      int j[ND];
      int r = e;
#pragma unroll
      for (int ax = ND - 1; ax >= 0; --ax) { j[ax] = r & 3; r >>= 2; }
      // value of this neighbourhood entry (and, JVP, of its tangent), ghost rows expanded
      T val = T(0), vdot = T(0);
      T wall = T(1);
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) wall *= tap[ax][j[ax]].v;
      const int c0 = src[0][j[0]].cnt, c1 = ND > 1 ? src[ND > 1 ? 1 : 0][j[ND > 1 ? 1 : 0]].cnt : 1;
      const int c2 = ND > 2 ? src[ND > 2 ? 2 : 0][j[ND > 2 ? 2 : 0]].cnt : 1;
      for (int k0 = 0; k0 < c0; ++k0)
        for (int k1 = 0; k1 < c1; ++k1)
          for (int k2 = 0; k2 < c2; ++k2) {
            const int kk[3] = {k0, k1, k2};
            int64_t idx = 0;
            T w = T(1);
#pragma unroll
            for (int ax = 0; ax < ND; ++ax) {
              const HaloSrc& h = src[ax][j[ax]];
              idx = idx * a.n[ax] + h.idx[kk[ax]];
              if (h.cnt == 2) w *= kk[ax] == 0 ? T(2) : T(-1);
            }
            val += w * a.f[idx];
            if (a.jvp) {
              if (a.f_dot != nullptr) vdot += w * a.f_dot[idx];
            } else if (a.grad_f != nullptr) {
              grad_atomic_add(a.grad_f + idx, cot * wall * w);
            }
          }
      if (a.jvp) tan += wall * vdot;
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) {
        T other = val;
#pragma unroll
        for (int bx = 0; bx < ND; ++bx)
          if (bx != ax) other *= tap[bx][j[bx]].v;
        part[ax][j[ax]] += other;
      }
    }
    // d out / d (local knot k of axis ax) = sum_j dT[ax][j]/dx_k * part[ax][j]
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      if (a.jvp ? a.knots_dot[ax] == nullptr : a.grad_knots[ax] == nullptr) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int p = cell[ax] - 2 + k;
        // (class-form periodic axes: all four local knots exist, they are original knots)
        if (a.kind[ax] != 1 && (p < 0 || p >= a.nk[ax])) continue;
        T g = T(0);
#pragma unroll
        for (int j = 0; j < 4; ++j) g += tap[ax][j].d[k] * part[ax][j];
        // lookup position -> index in the caller's knot vector
        int o = p;
        if (a.kind[ax] != 0) {
          o = (p - 1) % a.n[ax];
          if (o < 0) o += a.n[ax];
          if (a.perm[ax] != nullptr) o = a.perm[ax][o];
        }
        if (a.jvp) tan += g * a.knots_dot[ax][o];
        else grad_atomic_add(a.grad_knots[ax] + o, cot * g);
      }
    }
    if (a.jvp) a.tangent[qi] = tan;
  }
}

template <typename T, int ND>
static int launch_grad(const T* const* q, int64_t nq, const T* const* knots, const T* const* slope_knots,
                       const int32_t* const* perm,
                       const int32_t* n, const T* f, int32_t slope_method, double c, int32_t periodic_mask,
                       const ipx_params* params, int32_t params_on_device, const T* cot, T* const* grad_knots,
                       T* grad_f, const T* const* knots_dot, const T* f_dot, T* tangent, int jvp, void* stream) {
  if (nq < 0 || q == nullptr || knots == nullptr || n == nullptr || params == nullptr || f == nullptr)
    return IPX_EINVAL;
  if (slope_method != IPX_SLOPE_CUBIC && slope_method != IPX_SLOPE_CARDINAL) return IPX_EUNSUPPORTED;
  if (nq == 0) return IPX_OK;
  if (jvp ? tangent == nullptr : cot == nullptr) return IPX_EINVAL;
  GradArgs<T, ND> a;
  for (int ax = 0; ax < ND; ++ax) {
    if (q[ax] == nullptr || knots[ax] == nullptr || n[ax] < 2) return IPX_EINVAL;
    const int ring = (periodic_mask >> ax) & 1, padfirst = (periodic_mask >> (4 + ax)) & 1;
    if (ring && padfirst) return IPX_EINVAL;
    a.q[ax] = q[ax];
    a.knots[ax] = knots[ax];
    a.sknots[ax] = nullptr;
    if (ring) {
      // (original knots must be sorted inside one period: their neighbours are the table's)
      if (slope_knots == nullptr || slope_knots[ax] == nullptr || (perm != nullptr && perm[ax] != nullptr))
        return IPX_EUNSUPPORTED;
      a.sknots[ax] = slope_knots[ax];
    }
    a.perm[ax] = (ring || padfirst) && perm != nullptr ? perm[ax] : nullptr;
    a.n[ax] = n[ax];
    a.kind[ax] = ring ? 1 : (padfirst ? 2 : 0);
    a.nk[ax] = a.kind[ax] ? n[ax] + 2 : n[ax];
    a.grad_knots[ax] = grad_knots != nullptr ? grad_knots[ax] : nullptr;
    a.knots_dot[ax] = knots_dot != nullptr ? knots_dot[ax] : nullptr;
  }
  a.f = f;
  a.cot = cot;
  a.grad_f = grad_f;
  a.f_dot = f_dot;
  a.tangent = tangent;
  a.nq = nq;
  a.fac = T(1) - T(c);
  a.cardinal = slope_method == IPX_SLOPE_CARDINAL;
  a.jvp = jvp;
  if (params_on_device) {
    a.dp = params;
    a.hp = ipx_params{};
  } else {
    a.dp = nullptr;
    a.hp = *params;
  }
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (nq + kGradThreads - 1) / kGradThreads;
  int64_t cap = (int64_t)sms * 16;
  grad_stencil_kernel<T, ND><<<(int)(need < cap ? need : cap), kGradThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_grad_any(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,
                           const T* const* slope_knots, const int32_t* const* perm, const int32_t* n, const T* f, int32_t slope_method, double c,
                           int32_t periodic_mask, const ipx_params* params, int32_t pod, const T* cot,
                           T* const* grad_knots, T* grad_f, const T* const* knots_dot, const T* f_dot, T* tangent,
                           int jvp, void* stream) {
  if (ndim == 1)
    return launch_grad<T, 1>(q, nq, knots, slope_knots, perm, n, f, slope_method, c, periodic_mask, params, pod, cot, grad_knots,
                             grad_f, knots_dot, f_dot, tangent, jvp, stream);
  if (ndim == 2)
    return launch_grad<T, 2>(q, nq, knots, slope_knots, perm, n, f, slope_method, c, periodic_mask, params, pod, cot, grad_knots,
                             grad_f, knots_dot, f_dot, tangent, jvp, stream);
  if (ndim == 3)
    return launch_grad<T, 3>(q, nq, knots, slope_knots, perm, n, f, slope_method, c, periodic_mask, params, pod, cot, grad_knots,
                             grad_f, knots_dot, f_dot, tangent, jvp, stream);
  return IPX_EINVAL;
}

}  // namespace ipx

#define IPX_DEFINE_GRAD(SUF, T)                                                                                      \
  extern "C" int ipx_vjp_stencil_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,           \
                                       const T* const* slope_knots, const int32_t* const* perm, const int32_t* n, const T* f,                     \
                                       int32_t slope_method, double c, int32_t periodic_mask,                        \
                                       const ipx_params* params, int32_t params_on_device, const T* cotangent,       \
                                       T* const* grad_knots, T* grad_f, void* stream) {                              \
    return ipx::launch_grad_any<T>(ndim, q, nq, knots, slope_knots, perm, n, f, slope_method, c, periodic_mask, params,           \
                                   params_on_device, cotangent, grad_knots, grad_f, nullptr, nullptr, nullptr, 0,    \
                                   stream);                                                                          \
  }                                                                                                                  \
  extern "C" int ipx_jvp_stencil_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,           \
                                       const T* const* slope_knots, const int32_t* const* perm, const int32_t* n, const T* f,                     \
                                       int32_t slope_method, double c, int32_t periodic_mask,                        \
                                       const ipx_params* params, int32_t params_on_device,                           \
                                       const T* const* knots_dot, const T* f_dot, T* tangent, void* stream) {        \
    return ipx::launch_grad_any<T>(ndim, q, nq, knots, slope_knots, perm, n, f, slope_method, c, periodic_mask, params,           \
                                   params_on_device, nullptr, nullptr, nullptr, knots_dot, f_dot, tangent, 1,        \
                                   stream);                                                                          \
  }

IPX_DEFINE_GRAD(f32, float)
IPX_DEFINE_GRAD(f64, double)
