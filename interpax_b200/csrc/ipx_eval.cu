// ipx_eval.cu -- fused query evaluation for interp1d / interp2d / interp3d.
//
// One kernel does, per (query, batch element): periodic wrap of the query -> interval
// lookup on every axis -> local coordinate -> gather of the 2^d cell corners (x 2^d
// tables for the cubic family) -> separable Hermite / multilinear / nearest evaluation
// (value or derivative) -> extrapolation select -> one coalesced store.  Nothing but
// the queries, the touched table entries and the results moves through HBM.
//
// Replaces (reference repo paths): interpax/_spline.py:529-592 (1-D), :702-805 (2-D),
// :940-1105 (3-D), with _get_t_der :1138-1151 and _extrap :1180-1222 folded in.
//
// The reference contracts a dense 4^d x 4^d matrix (A_CUBIC / A_BICUBIC / A_TRICUBIC)
// with the gathered values.  Those matrices are tensor powers of A_CUBIC
// (tests/test_oracle_pins.py::test_separable_form_equals_dense_form), so the kernels
// evaluate the separable form: per-axis Hermite weights, then one axis at a time
// (64 + 16 + 4 FMAs in 3-D instead of 4096 + 64).
#include <cuda_runtime.h>

#include <atomic>

#include "ipx_contract.cuh"
#include "ipx_device.cuh"
#include "ipx_internal.cuh"

#define IPX_STR2(x) #x
#define IPX_STR(x) IPX_STR2(x)

namespace ipx {

thread_local RunIf g_run_if = {nullptr, 0u};
static std::atomic<unsigned long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

constexpr int kThreads = 256;

template <typename T, int ND> struct EvalArgs {
  AxisDesc<T> ax[ND];
  const T* __restrict__ tab[1 << ND];  // canonical record order r (see record_slot)
  T* __restrict__ out;
  int32_t* __restrict__ idx_out;
  int64_t nq;
  int64_t batch;
  ipx_params hp;          // host copy of the traced operands ...
  const ipx_params* dp;   // ... or a device pointer to them (takes precedence)
  const int32_t* __restrict__ run_flag;  // NULL, or: run only if bit *run_flag of run_accept is set
  uint32_t run_accept;                   // (ipx_internal.cuh)
};

template <typename T, int ND> __device__ __forceinline__ bool skip_launch(const EvalArgs<T, ND>& a) {
  return a.run_flag != nullptr && ((a.run_accept >> (*a.run_flag & 31)) & 1u) == 0;
}

// linear node index of corner (c0[,c1[,c2]]) times batch
template <typename T, int ND>
__device__ __forceinline__ int64_t node_index(const EvalArgs<T, ND>& a, const AxisLoc<T>* L,
                                              int corner_bits) {
  // corner bit a (bit ND-1-axis ... ) : we use bit `ax` = corner along axis ax
  int64_t idx = (corner_bits & 1) ? L[0].t1 : L[0].t0;
#pragma unroll
  for (int ax = 1; ax < ND; ++ax) {
    int t = ((corner_bits >> ax) & 1) ? L[ax].t1 : L[ax].t0;
    idx = idx * a.ax[ax].n + t;
  }
  return idx;
}

// Load the 2^ND values of one (node, batch element) in canonical record order:
// bit (ND-1-ax) of r set  <=>  the table is differentiated along axis ax, i.e.
//   1-D: [f, fx]   2-D: [f, fy, fx, fxy]   3-D: [f, fz, fy, fyz, fx, fxz, fxy, fxyz]
template <typename T, int ND, int LAYOUT>
__device__ __forceinline__ void load_record(const EvalArgs<T, ND>& a, int64_t rec_index, T* r) {
  if (LAYOUT == IPX_AOS) {
    const T* p = a.tab[0] + rec_index * (1 << ND);
    if (ND == 1) ldg_rec2(p, r);
    if (ND == 2) ldg_rec4(p, r);
    if (ND == 3) ldg_rec8(p, r);
  } else {
#pragma unroll
    for (int k = 0; k < (1 << ND); ++k) r[k] = ldg(a.tab[k] + rec_index);
  }
}

template <typename T, int ND>
__device__ __forceinline__ T apply_extrap(const EvalArgs<T, ND>& a, const ipx_params& P,
                                          const AxisLoc<T>* L, T v) {
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    if (a.ax[ax].wrap) continue;  // lowx = highx = True, _spline.py:527
    const ipx_axis_params& p = P.axis[ax];
    if (!p.keep_lo && L[ax].below) v = T(p.fill_lo);
    if (!p.keep_hi && L[ax].above) v = T(p.fill_hi);
  }
  return v;
}

// ------------------------------------------------------------------------------------
// cubic family: weights once per query, then per batch element two node layers
// ------------------------------------------------------------------------------------
template <typename T, int ND, int LAYOUT>
__device__ __forceinline__ T eval_cubic(const EvalArgs<T, ND>& a, const AxisLoc<T>* L,
                                        const T (*w)[4], int64_t b) {
  WeightRefs<T, ND, 1> W;
  W.w[0] = w;
  T res = T(0);
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    T A[1][2];
    layer_contract<T, ND, 1>(
        [&](int c, T* r) {
          load_record<T, ND, LAYOUT>(a, node_index<T, ND>(a, L, c | (kk << (ND - 1))) * a.batch + b, r);
        },
        W, A);
    res = fma(A[0][0], w[ND - 1][kk], fma(A[0][1], w[ND - 1][2 + kk], res));
  }
  return res;
}

// ------------------------------------------------------------------------------------
// linear
// ------------------------------------------------------------------------------------
template <typename T, int ND> struct LinearW {
  T w[ND][2];
  T scale;
};

template <typename T, int ND>
__device__ __forceinline__ void linear_setup(const ipx_params& P, const AxisLoc<T>* L, LinearW<T, ND>& W) {
  W.scale = T(1);
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    T dxi;
    linear_weights(L[ax], P.axis[ax].derivative, W.w[ax], dxi);
    W.scale = ax == 0 ? dxi : W.scale * dxi;
  }
}

template <typename T, int ND>
__device__ __forceinline__ T eval_linear(const EvalArgs<T, ND>& a, const ipx_params& P,
                                         const AxisLoc<T>* L, const LinearW<T, ND>& W, int64_t b) {
  const T* __restrict__ f = a.tab[0];
  if (ND == 1) {
    // _spline.py:542-567
    int order = P.axis[0].derivative;
    if (order >= 2) return T(0);
    T f0 = ldg(f + (int64_t)L[0].t0 * a.batch + b);
    T f1 = ldg(f + (int64_t)L[0].t1 * a.batch + b);
    T df = f1 - f0;
    T dx = L[0].x1 - L[0].x0;
    T dxi = safe_recip(dx);
    if (order == 1) return df * dxi;
    T delta = L[0].q - L[0].x0;
    return dx == T(0) ? f1 : f0 + (delta * dxi) * df;
  }
  // _spline.py:728-755, 982-1023:  dxi*dyi[*dzi] * sum F * tx * ty [* tz]
  T s = T(0);
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) {
    T val = ldg(f + node_index<T, ND>(a, L, c) * a.batch + b);
    T wt = W.w[0][c & 1];
#pragma unroll
    for (int ax = 1; ax < ND; ++ax) wt *= W.w[ax][(c >> ax) & 1];
    s += val * wt;
  }
  return W.scale * s;
}

// ------------------------------------------------------------------------------------
// nearest: returns the node index (times batch not applied) of the selected knot/corner,
// or -1 when the result is identically zero (derivative requested)
// ------------------------------------------------------------------------------------
template <typename T, int ND>
__device__ __forceinline__ int64_t nearest_node(const EvalArgs<T, ND>& a, const ipx_params& P,
                                                const AxisLoc<T>* L) {
  if (ND == 1) {
    // _spline.py:531-538: argmin_k |q - x[k]| over ALL knots, first minimum wins.
    if (P.axis[0].derivative >= 1) return -1;
    const AxisDesc<T>& A = a.ax[0];
    T q = L[0].q;
    int c;
    if (q != q) {
      c = 0;  // every distance is NaN: argmin returns the first
    } else {
      T d0 = t_abs(q - L[0].x0), d1 = t_abs(q - L[0].x1);
      c = (d0 <= d1) ? L[0].i - 1 : L[0].i;
      T dmin = (d0 <= d1) ? d0 : d1;
      // |q - x[k]| is non-increasing in k up to the nearest knot, also after rounding,
      // so ties (duplicate knots, or distances that round together) form a contiguous
      // run ending at c: find its first element.
      if (c > 0 && t_abs(q - A.knots[c - 1]) <= dmin) {
        int lo = 0, hi = c - 1;  // predicate true at hi
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (t_abs(q - A.knots[mid]) <= dmin) hi = mid; else lo = mid + 1;
        }
        c = lo;
      }
    }
    return padded_to_table(A, c);
  }
  // _spline.py:704-726, 942-980: nearest (Euclidean) of the cell corners; the corner
  // order puts the UPPER knot first on every axis, x fastest; first minimum wins.
  bool all_zero = true;
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) all_zero = all_zero && (P.axis[ax].derivative == 0);
  if (!all_zero) return -1;
  T best = T(0);
  int best_c = 0;
#pragma unroll
  for (int c = 0; c < (1 << ND); ++c) {
    T s = T(0);
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      T xk = ((c >> ax) & 1) ? L[ax].x0 : L[ax].x1;
      T d = t_abs(xk - L[ax].q);
      s = ax == 0 ? mul_rn(d, d) : add_rn(s, mul_rn(d, d));
    }
    T dist = t_sqrt(s);
    // jnp.argmin: first minimum; a NaN distance wins over any number and the first NaN stays
    bool take = (c == 0) || (!(best != best) && ((dist != dist) || dist < best));
    if (take) { best = dist; best_c = c; }
  }
  // corner bit set = LOWER knot here; node_index wants bit set = upper
  return node_index<T, ND>(a, L, (~best_c) & ((1 << ND) - 1));
}

// ------------------------------------------------------------------------------------
// the generic kernel.  Lane mapping: a query owns 2^bpad_log2 adjacent lanes which stride over
// the batch (trailing dims of f), so the search and the weights are computed once per lane and
// reused for its batch elements, table reads are contiguous in the batch index and the
// (nq, batch) result is written coalesced.  batch == 1 degenerates to one lane per query.
// ------------------------------------------------------------------------------------
// Queries a lane handles per round.  Batched evaluation is a dependent chain per query
// (search -> gather -> store) with little arithmetic, so several independent queries are
// kept in flight per lane to cover the memory latency; in 3-D the register cost is too high.
template <int ND> struct QueriesPerRound {
  static constexpr int value = ND == 1 ? 4 : (ND == 2 ? 2 : 1);
};

template <typename T, int ND, int METHOD, int LAYOUT>
__global__ void __launch_bounds__(kThreads) eval_kernel(const EvalArgs<T, ND> a, int bpad_log2) {
  constexpr int U = QueriesPerRound<ND>::value;
  __shared__ AxisConst<T> consts[ND];
  __shared__ ipx_params params_s;
  if (skip_launch(a)) return;
  if (threadIdx.x < ND) axis_consts(a.ax[threadIdx.x], consts[threadIdx.x]);
  if (threadIdx.x == 32) params_s = a.dp ? *a.dp : a.hp;
  __syncthreads();
  const ipx_params& P = params_s;

  const int lane = threadIdx.x & 31;
  const int bpad = 1 << bpad_log2;
  const int qpw = 32 >> bpad_log2;  // queries per warp per sub-round
  const int qsub = lane >> bpad_log2;
  const int b0 = lane & (bpad - 1);
  const int64_t warps = (int64_t)gridDim.x * (kThreads / 32);
  const int64_t wg = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (b0 >= a.batch) return;  // padding lanes (no warp-level sync below)

  for (int64_t qbase = wg * qpw * U; qbase < a.nq; qbase += warps * qpw * U) {
    AxisLoc<T> L[U][ND];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t qi = qbase + u * qpw + qsub;
      valid[u] = qi < a.nq;
      const int64_t qs = valid[u] ? qi : 0;  // harmless in-range index for the idle slots
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) locate(a.ax[ax], consts[ax], P.axis[ax], qs, L[u][ax]);
    }

    if (METHOD == IPX_CUBIC) {
      T w[U][ND][4];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) hermite_weights(L[u][ax], P.axis[ax].derivative, w[u][ax]);
      for (int64_t b = b0; b < a.batch; b += bpad) {
        T v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = eval_cubic<T, ND, LAYOUT>(a, L[u], w[u], b);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (valid[u]) a.out[(qbase + u * qpw + qsub) * a.batch + b] = apply_extrap<T, ND>(a, P, L[u], v[u]);
      }
    } else if (METHOD == IPX_LINEAR) {
      LinearW<T, ND> W[U];
      if (ND > 1) {
#pragma unroll
        for (int u = 0; u < U; ++u) linear_setup<T, ND>(P, L[u], W[u]);
      }
      for (int64_t b = b0; b < a.batch; b += bpad) {
        T v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = eval_linear<T, ND>(a, P, L[u], W[u], b);
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (valid[u]) a.out[(qbase + u * qpw + qsub) * a.batch + b] = apply_extrap<T, ND>(a, P, L[u], v[u]);
      }
    } else {
      int64_t node[U];
#pragma unroll
      for (int u = 0; u < U; ++u) node[u] = nearest_node<T, ND>(a, P, L[u]);
      for (int64_t b = b0; b < a.batch; b += bpad) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          T v = node[u] < 0 ? T(0) : ldg(a.tab[0] + node[u] * a.batch + b);
          if (valid[u]) a.out[(qbase + u * qpw + qsub) * a.batch + b] = apply_extrap<T, ND>(a, P, L[u], v);
        }
      }
    }
    if (a.idx_out != nullptr && b0 == 0) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!valid[u]) continue;
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) a.idx_out[(int64_t)ax * a.nq + qbase + u * qpw + qsub] = L[u][ax].i;
      }
    }
  }
}

}  // namespace ipx

#include "ipx_eval_tile.cuh"
#include "ipx_eval_rows.cuh"

namespace ipx {

template <typename T> __global__ void bin_index_kernel(const T* __restrict__ knots, int n,
                                                       const T* __restrict__ q, int64_t nq,
                                                       int32_t* __restrict__ idx) {
  AxisDesc<T> A{q, knots, nullptr, n, n, 0, 0};
  AxisConst<T> C;
  axis_consts(A, C);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nq; e += stride) {
    T lo, hi;
    idx[e] = bin_index(knots, n, q[e], C.x_first, C.inv_h, lo, hi);
  }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
static int grid_for(int64_t total, int threads) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (total + threads - 1) / threads;
  int64_t cap = (int64_t)sms * 32;  // a multiple of the SM count; grid-stride beyond it
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// reference stacking order -> canonical record slot (bit ND-1-ax <=> d/d(ax))
static int record_slot(int ndim, int ff) {
  // which axes the reference's ff-th table is differentiated along (x=1, y=2, z=4)
  static const int d1[2] = {0, 1};
  static const int d2[4] = {0, 1, 2, 3};               // f, fx, fy, fxy
  static const int d3[8] = {0, 1, 2, 4, 3, 5, 6, 7};   // f, fx, fy, fz, fxy, fxz, fyz, fxyz
  int mask = ndim == 1 ? d1[ff] : ndim == 2 ? d2[ff] : d3[ff];
  int slot = 0;
  for (int ax = 0; ax < ndim; ++ax)
    if (mask & (1 << ax)) slot |= 1 << (ndim - 1 - ax);
  return slot;
}

template <typename T, int ND>
static int launch_eval(const T* const* q, int64_t nq, const T* const* knots,
                       const int32_t* const* perm, const int32_t* n, const T* const* tables,
                       int32_t layout, int64_t batch, int32_t method, int32_t periodic_mask,
                       const ipx_params* params, int32_t params_on_device, T* out,
                       int32_t* idx_out, void* stream) {
  if (nq < 0 || batch < 1 || tables == nullptr || params == nullptr) return IPX_EINVAL;
  if (method != IPX_NEAREST && method != IPX_LINEAR && method != IPX_CUBIC) return IPX_EINVAL;
  if (layout != IPX_SOA && layout != IPX_AOS) return IPX_EINVAL;
  if (layout == IPX_AOS && method != IPX_CUBIC) return IPX_EUNSUPPORTED;
  if (nq == 0) return IPX_OK;
  if (out == nullptr) return IPX_EINVAL;

  EvalArgs<T, ND> a;
  for (int ax = 0; ax < ND; ++ax) {
    if (q[ax] == nullptr || knots[ax] == nullptr || n[ax] < 2) return IPX_EINVAL;
    int per = (periodic_mask >> ax) & 1;
    int wrap_only = (periodic_mask >> (4 + ax)) & 1;
    if (per && wrap_only) return IPX_EINVAL;
    a.ax[ax].q = q[ax];
    a.ax[ax].knots = knots[ax];
    a.ax[ax].perm = per ? perm[ax] : nullptr;
    a.ax[ax].n = n[ax];
    a.ax[ax].nk = per ? n[ax] + 2 : n[ax];
    a.ax[ax].periodic = per;
    a.ax[ax].wrap = per | wrap_only;
  }
  for (int k = 0; k < (1 << ND); ++k) a.tab[k] = nullptr;
  if (method == IPX_CUBIC && layout == IPX_SOA) {
    for (int ff = 0; ff < (1 << ND); ++ff) {
      if (tables[ff] == nullptr) return IPX_EINVAL;
      a.tab[record_slot(ND, ff)] = tables[ff];
    }
  } else {
    if (tables[0] == nullptr) return IPX_EINVAL;
    a.tab[0] = tables[0];
  }
  a.out = out;
  a.idx_out = idx_out;
  a.nq = nq;
  a.batch = batch;
  if (params_on_device) {
    a.dp = params;
    a.hp = ipx_params{};
  } else {
    a.dp = nullptr;
    a.hp = *params;
  }
  a.run_flag = g_run_if.flag;
  a.run_accept = g_run_if.accept;
  g_run_if = RunIf{nullptr, 0u};

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int bpad_log2 = 0;  // lanes per query = smallest power of two >= min(batch, 32)
  while ((1 << bpad_log2) < 32 && (int64_t)(1 << bpad_log2) < batch) ++bpad_log2;
  int grid = grid_for(((nq << bpad_log2) + QueriesPerRound<ND>::value - 1) / QueriesPerRound<ND>::value, kThreads);
  if constexpr (ND == 1) {
    // wide vector-valued 1-D tables: the row-streaming kernel
    if (launch_rows<T>(a, method, layout, s)) {
      count_launches(1);
      return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
    }
  }
  if (method == IPX_CUBIC) {
    if (layout == IPX_AOS) {
      // scalar-valued packed tables: the warp-cooperative shared-memory kernel
      const bool tiled = batch == 1 && launch_tile<T, ND>(a, s);
      if (!tiled) eval_kernel<T, ND, IPX_CUBIC, IPX_AOS><<<grid, kThreads, 0, s>>>(a, bpad_log2);
    } else eval_kernel<T, ND, IPX_CUBIC, IPX_SOA><<<grid, kThreads, 0, s>>>(a, bpad_log2);
  } else if (method == IPX_LINEAR) {
    eval_kernel<T, ND, IPX_LINEAR, IPX_SOA><<<grid, kThreads, 0, s>>>(a, bpad_log2);
  } else {
    eval_kernel<T, ND, IPX_NEAREST, IPX_SOA><<<grid, kThreads, 0, s>>>(a, bpad_log2);
  }
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_bin_index(const T* knots, int32_t n, const T* q, int64_t nq, int32_t* idx,
                            void* stream) {
  if (knots == nullptr || n < 2 || nq < 0) return IPX_EINVAL;
  if (nq == 0) return IPX_OK;
  if (q == nullptr || idx == nullptr) return IPX_EINVAL;
  bin_index_kernel<T><<<grid_for(nq, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      knots, n, q, nq, idx);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

}  // namespace ipx

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
#define IPX_DEFINE_EVAL(SUF, T)                                                                  \
  extern "C" int ipx_eval1d_##SUF(const T* xq, int64_t nq, const T* x, const int32_t* permx,     \
                                  int32_t nx, const T* const* tables, int32_t layout,            \
                                  int64_t batch, int32_t method_class, int32_t periodic_mask,    \
                                  const ipx_params* params, int32_t params_on_device, T* out,    \
                                  int32_t* idx_out, void* stream) {                              \
    const T* q[1] = {xq};                                                                        \
    const T* k[1] = {x};                                                                         \
    const int32_t* p[1] = {permx};                                                               \
    int32_t n[1] = {nx};                                                                         \
    return ipx::launch_eval<T, 1>(q, nq, k, p, n, tables, layout, batch, method_class,           \
                                  periodic_mask, params, params_on_device, out, idx_out, stream);\
  }                                                                                              \
  extern "C" int ipx_eval2d_##SUF(const T* xq, const T* yq, int64_t nq, const T* x, const T* y,  \
                                  const int32_t* permx, const int32_t* permy, int32_t nx,        \
                                  int32_t ny, const T* const* tables, int32_t layout,            \
                                  int64_t batch, int32_t method_class, int32_t periodic_mask,    \
                                  const ipx_params* params, int32_t params_on_device, T* out,    \
                                  int32_t* idx_out, void* stream) {                              \
    const T* q[2] = {xq, yq};                                                                    \
    const T* k[2] = {x, y};                                                                      \
    const int32_t* p[2] = {permx, permy};                                                        \
    int32_t n[2] = {nx, ny};                                                                     \
    return ipx::launch_eval<T, 2>(q, nq, k, p, n, tables, layout, batch, method_class,           \
                                  periodic_mask, params, params_on_device, out, idx_out, stream);\
  }                                                                                              \
  extern "C" int ipx_eval3d_##SUF(const T* xq, const T* yq, const T* zq, int64_t nq, const T* x, \
                                  const T* y, const T* z, const int32_t* permx,                  \
                                  const int32_t* permy, const int32_t* permz, int32_t nx,        \
                                  int32_t ny, int32_t nz, const T* const* tables,                \
                                  int32_t layout, int64_t batch, int32_t method_class,           \
                                  int32_t periodic_mask, const ipx_params* params,               \
                                  int32_t params_on_device, T* out, int32_t* idx_out,            \
                                  void* stream) {                                                \
    const T* q[3] = {xq, yq, zq};                                                                \
    const T* k[3] = {x, y, z};                                                                   \
    const int32_t* p[3] = {permx, permy, permz};                                                 \
    int32_t n[3] = {nx, ny, nz};                                                                 \
    return ipx::launch_eval<T, 3>(q, nq, k, p, n, tables, layout, batch, method_class,           \
                                  periodic_mask, params, params_on_device, out, idx_out, stream);\
  }

IPX_DEFINE_EVAL(f32, float)
IPX_DEFINE_EVAL(f64, double)

extern "C" int ipx_bin_index_f32(const float* knots, int32_t n, const float* q, int64_t nq,
                                 int32_t* idx, void* stream) {
  return ipx::launch_bin_index<float>(knots, n, q, nq, idx, stream);
}
extern "C" int ipx_bin_index_f64(const double* knots, int32_t n, const double* q, int64_t nq,
                                 int32_t* idx, void* stream) {
  return ipx::launch_bin_index<double>(knots, n, q, nq, idx, stream);
}

extern "C" int ipx_version(void) { return IPX_VERSION; }
extern "C" uint64_t ipx_launch_count(void) { return ipx::g_launches.load(std::memory_order_relaxed); }
extern "C" const char* ipx_build_info(void) {
  return "libipx " __DATE__ " nvcc " IPX_STR(__CUDACC_VER_MAJOR__) "." IPX_STR(
      __CUDACC_VER_MINOR__) " arch sm_100a";
}
