// ipx_eval_rows.cuh -- 1-D evaluation of wide vector-valued tables (f of shape (Nx, B), B >= 32):
// the C2 shape of BASELINE.json (1e6 knots, B = 64, 1e9 queries).  Included by ipx_eval.cu.
//
// The result write (B values per query) is 98 % of the algorithmic bytes, so the kernel must
// stream it at HBM speed and keep everything else off the memory system:
//   * a warp takes 32 consecutive queries; lane j locates query j (one search per query, not
//     one per (query, column) as a flat element mapping would);
//   * the warp then walks its queries in order with lanes mapped to COLUMNS: the located
//     interval and the four Hermite weights of query j are broadcast by shuffles, every lane
//     evaluates two columns and the warp writes 2 x 256 contiguous bytes of the result row;
//   * the two table rows of the interval are kept in registers and re-read only when the
//     interval changes (a warp-uniform branch), so sorted or clustered queries read each table
//     row about once -- random queries still cost two row reads per query, which is DRAM
//     traffic no kernel can avoid.
// Arithmetic: the same separable form as everywhere else (ipx_contract.cuh, ND == 1):
//   res = fma(f1, w1, fma(fx1, w3, fma(f0, w0, fma(fx0, w2, 0)))).
#pragma once

namespace ipx {

constexpr int kRowThreads = 256;

template <typename T, int LAYOUT>
__device__ __forceinline__ void load_row_pair(const EvalArgs<T, 1>& a, int64_t rec, T& f, T& fx) {
  if (LAYOUT == IPX_AOS) {
    T r[2];
    ldg_rec2(a.tab[0] + rec * 2, r);
    f = r[0];
    fx = r[1];
  } else {
    f = ldg(a.tab[0] + rec);
    fx = ldg(a.tab[1] + rec);
  }
}

template <typename T> __device__ __forceinline__ T shfl_t(T v, int src);
template <> __device__ __forceinline__ float shfl_t<float>(float v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}
template <> __device__ __forceinline__ double shfl_t<double>(double v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}

// METHOD: IPX_CUBIC (tables f, fx; AoS or SoA) or IPX_LINEAR (table f)
template <typename T, int METHOD, int LAYOUT>
__global__ void __launch_bounds__(kRowThreads) eval1d_rows_kernel(const EvalArgs<T, 1> a) {
  if (skip_launch(a)) return;
  __shared__ AxisConst<T> consts;
  __shared__ ipx_params params_s;
  if (threadIdx.x == 0) axis_consts(a.ax[0], consts);
  if (threadIdx.x == 32) params_s = a.dp ? *a.dp : a.hp;
  __syncthreads();
  const ipx_axis_params& P = params_s.axis[0];
  const AxisDesc<T>& A = a.ax[0];

  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (kRowThreads / 32);
  const int64_t wg = (int64_t)blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const int64_t B = a.batch;

  for (int64_t qbase = wg * 32; qbase < a.nq; qbase += warps * 32) {
    // ---- phase 1: lane j locates query qbase + j and forms its per-query scalars ----
    const int64_t qi = qbase + lane;
    AxisLoc<T> L;
    locate(A, consts, P, qi < a.nq ? qi : qbase, L);
    T c[4];
    if (METHOD == IPX_CUBIC) {
      hermite_weights(L, P.derivative, c);
    } else {
      // _spline.py:542-567: order 0  f0 + (delta*dxi)*df, or f1 where dx == 0; order 1  df*dxi
      const T dx = L.x1 - L.x0;
      const T dxi = safe_recip(dx);
      c[0] = (L.q - L.x0) * dxi;
      c[1] = dxi;
      c[2] = dx == T(0) ? T(1) : T(0);
      c[3] = T(0);
    }
    int fill = 0;  // 1: lower fill applies, 2: upper fill applies (upper wins, as in _extrap)
    if (!A.wrap) {
      if (!P.keep_lo && L.below) fill |= 1;
      if (!P.keep_hi && L.above) fill |= 2;
    }
    if (a.idx_out != nullptr && qi < a.nq) a.idx_out[qi] = L.i;
    const int nv = (int)min((int64_t)32, a.nq - qbase);

    // ---- phase 2: lanes are columns; walk the warp's queries ----
    for (int64_t bb = 0; bb < B; bb += 64) {
      const int64_t b0 = bb + lane, b1 = bb + 32 + lane;
      const bool v0 = b0 < B, v1 = b1 < B;
      int cur0 = -1, cur1 = -1;
      T f00 = T(0), g00 = T(0), f01 = T(0), g01 = T(0);  // column b0: node t0 (f, fx), node t1
      T f10 = T(0), g10 = T(0), f11 = T(0), g11 = T(0);  // column b1
      for (int j = 0; j < nv; ++j) {
        const int t0 = __shfl_sync(0xffffffffu, L.t0, j);
        const int t1 = __shfl_sync(0xffffffffu, L.t1, j);
        const T w0 = shfl_t(c[0], j), w1 = shfl_t(c[1], j), w2 = shfl_t(c[2], j), w3 = shfl_t(c[3], j);
        const int fl = __shfl_sync(0xffffffffu, fill, j);
        if (t0 != cur0 || t1 != cur1) {  // warp-uniform: the interval changed, fetch its rows
          cur0 = t0;
          cur1 = t1;
          if (METHOD == IPX_CUBIC) {
            if (v0) { load_row_pair<T, LAYOUT>(a, (int64_t)t0 * B + b0, f00, g00); load_row_pair<T, LAYOUT>(a, (int64_t)t1 * B + b0, f01, g01); }
            if (v1) { load_row_pair<T, LAYOUT>(a, (int64_t)t0 * B + b1, f10, g10); load_row_pair<T, LAYOUT>(a, (int64_t)t1 * B + b1, f11, g11); }
          } else {
            if (v0) { f00 = ldg(a.tab[0] + (int64_t)t0 * B + b0); f01 = ldg(a.tab[0] + (int64_t)t1 * B + b0); }
            if (v1) { f10 = ldg(a.tab[0] + (int64_t)t0 * B + b1); f11 = ldg(a.tab[0] + (int64_t)t1 * B + b1); }
          }
        }
        T r0, r1;
        if (METHOD == IPX_CUBIC) {
          r0 = fma(f01, w1, fma(g01, w3, fma(f00, w0, fma(g00, w2, T(0)))));
          r1 = fma(f11, w1, fma(g11, w3, fma(f10, w0, fma(g10, w2, T(0)))));
        } else {
          const int order = P.derivative;
          if (order >= 2) {
            r0 = T(0); r1 = T(0);
          } else if (order == 1) {
            r0 = (f01 - f00) * w1; r1 = (f11 - f10) * w1;
          } else {
            r0 = w2 != T(0) ? f01 : f00 + w0 * (f01 - f00);
            r1 = w2 != T(0) ? f11 : f10 + w0 * (f11 - f10);
          }
        }
        if (fl & 1) { r0 = T(P.fill_lo); r1 = r0; }
        if (fl & 2) { r0 = T(P.fill_hi); r1 = r0; }
        T* __restrict__ o = a.out + (qbase + j) * B;
        if (v0) o[b0] = r0;
        if (v1) o[b1] = r1;
      }
    }
  }
}

// applies to 1-D cubic / linear evaluation with at least 32 columns
template <typename T>
static bool launch_rows(const EvalArgs<T, 1>& a, int method, int layout, cudaStream_t s) {
  if (a.batch < 32 || (method != IPX_CUBIC && method != IPX_LINEAR)) return false;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (a.nq + 32 * (kRowThreads / 32) - 1) / (32 * (kRowThreads / 32));
  int64_t grid = (int64_t)sms * 8;
  if (need < grid) grid = need;
  if (grid < 1) grid = 1;
  if (method == IPX_LINEAR) eval1d_rows_kernel<T, IPX_LINEAR, IPX_SOA><<<(int)grid, kRowThreads, 0, s>>>(a);
  else if (layout == IPX_AOS) eval1d_rows_kernel<T, IPX_CUBIC, IPX_AOS><<<(int)grid, kRowThreads, 0, s>>>(a);
  else eval1d_rows_kernel<T, IPX_CUBIC, IPX_SOA><<<(int)grid, kRowThreads, 0, s>>>(a);
  return true;
}

}  // namespace ipx
