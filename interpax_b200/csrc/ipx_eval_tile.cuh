// ipx_eval_tile.cuh -- warp-cooperative, shared-memory-staged evaluation of the cubic family
// on AoS-packed tables (the Interpolator1D/2D/3D path).  Included by ipx_eval.cu.
//
// Why: a cubic query needs 2^d corner records (8 x 64 B in 3-D f64).  When every lane gathers
// its own records, each 128-bit load instruction of a warp touches one cache line per distinct
// cell in the warp and 32 such instructions are needed, so the L1 / LSU pipe, not HBM, bounds
// the kernel (profiles/ncu_eval3d_r01_baseline.txt).  Spatially coherent (cell-sorted) queries
// put a warp's 32 queries into a handful of consecutive cells along the fastest table axis.
// The warp therefore
//   1. groups its lanes by the cell "column" (interval indices on all axes but the last),
//   2. cooperatively copies the run of nodes the group needs -- 2^(d-1) columns x (run+1)
//      nodes, contiguous along the last axis -- global -> shared with full-line 128-bit
//      accesses (each byte crosses L1 once per warp instead of once per lane),
//   3. lets every lane read its 2^d records from shared memory (same-cell lanes broadcast),
//      and run the separable Hermite contraction.
// Incoherent queries (many groups in a warp) take the direct per-lane gather instead, which
// has more memory-level parallelism; both paths compute the same arithmetic.
//
// Knot vectors and reciprocal cell widths are staged in shared memory once per CTA (the
// grid is persistent), so the interval lookup and the local coordinate cost no global loads
// and no divisions per query.  1/dx is computed by the same IEEE division the direct path
// uses, so results are bit-identical between the two.
#pragma once

namespace ipx {

constexpr int kTileThreads = 256;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kRunCap = 16;       // nodes of one run held per column
constexpr int kMaxGroups = 4;     // more distinct columns than this in a warp -> direct path
constexpr int kSmemKnotCap = 3072;  // total knots (all axes) staged in shared memory

template <typename T, int ND> struct TileCfg {
  static constexpr int NT = 1 << ND;                       // values per record
  static constexpr int NC = 1 << (ND - 1);                 // columns per group
  static constexpr int REC_BYTES = NT * (int)sizeof(T);
  static constexpr int CH = REC_BYTES / 16;                // 16-byte chunks per record
  static constexpr int STRIDE = REC_BYTES + (REC_BYTES >= 32 ? 16 : 0);  // bank-conflict padding
  static constexpr int WARP_TILE_BYTES = NC * kRunCap * STRIDE;
};

// wrap for a positive period with fast exact paths around the principal interval:
//   [0,P): q     [-P,0): q+P (what fmod + sign fix gives)     [P,2P]: q-P (exact, Sterbenz)
template <typename T> __device__ __forceinline__ T py_mod_fast(T q, T P) {
  if (q >= T(0)) {
    if (q < P) return q;
    if (q <= P + P) return q - P;
  } else if (q >= -P) {
    T r = q + P;  // fmod(q,P) == q here, then the sign fix adds P
    return r;
  }
  return py_mod(q, P);
}

template <typename T> struct AxisSm {
  const T* knots;  // shared or global
  const T* rdx;    // shared reciprocal widths (rdx[i] = 1/(x[i]-x[i-1]) guarded) or nullptr
};

template <typename T>
__device__ __forceinline__ void locate_s(const AxisDesc<T>& A, const AxisSm<T>& S,
                                         const AxisConst<T>& C, const ipx_axis_params& P, T qraw,
                                         AxisLoc<T>& L, T& dxi) {
  T q = qraw;
  if (A.wrap) q = py_mod_fast(q, T(fabs(P.period)));
  L.q = q;
  L.i = bin_index(S.knots, A.nk, q, C.x_first, C.inv_h, L.x0, L.x1);
  L.below = q < C.x_first;
  L.above = q > C.x_last;
  dxi = S.rdx ? S.rdx[L.i] : safe_recip(L.x1 - L.x0);
  if (A.periodic) {
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

// hermite_weights with the reciprocal width supplied
template <typename T>
__device__ __forceinline__ void hermite_weights_r(const AxisLoc<T>& L, int order, T dxi, T w[4]) {
  T dx = L.x1 - L.x0;
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
  w[0] = tt0 - T(3) * tt2 + T(2) * tt3;
  w[1] = T(3) * tt2 - T(2) * tt3;
  w[2] = (tt1 - T(2) * tt2 + tt3) * dx;
  w[3] = (tt3 - tt2) * dx;
}

// 16-byte shared-memory accesses
template <typename T, int NT> __device__ __forceinline__ void lds_record(const char* p, T* r) {
#pragma unroll
  for (int c = 0; c < (NT * (int)sizeof(T)) / 16; ++c) {
    float4 v = *reinterpret_cast<const float4*>(p + 16 * c);
    const T* s = reinterpret_cast<const T*>(&v);
#pragma unroll
    for (int k = 0; k < 16 / (int)sizeof(T); ++k) r[c * (16 / (int)sizeof(T)) + k] = s[k];
  }
}

// separable contraction with a record loader: load(corner_bits, r) fills the 2^ND values of
// the node at corner bits (bit ax = upper corner along axis ax), canonical record order.
template <typename T, int ND, typename Loader>
__device__ __forceinline__ T contract_cubic(const T (*w)[4], Loader load) {
  T res = T(0);
  if (ND == 1) {
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      T r[2];
      load(ii, r);
      res += r[0] * w[0][ii] + r[1] * w[0][2 + ii];
    }
  } else if (ND == 2) {
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      T p[2] = {T(0), T(0)};
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        T r[4];
        load(ii | (jj << 1), r);
        p[0] += r[0] * w[1][jj] + r[1] * w[1][2 + jj];
        p[1] += r[2] * w[1][jj] + r[3] * w[1][2 + jj];
      }
      res += p[0] * w[0][ii] + p[1] * w[0][2 + ii];
    }
  } else {
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      T v[2] = {T(0), T(0)};
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        T p[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          T r[8];
          load(ii | (jj << 1) | (kk << 2), r);
#pragma unroll
          for (int m = 0; m < 4; ++m) p[m] += r[2 * m] * w[2][kk] + r[2 * m + 1] * w[2][2 + kk];
        }
        v[0] += p[0] * w[1][jj] + p[1] * w[1][2 + jj];
        v[1] += p[2] * w[1][jj] + p[3] * w[1][2 + jj];
      }
      res += v[0] * w[0][ii] + v[1] * w[0][2 + ii];
    }
  }
  return res;
}

// table index of knot position p along the LAST axis (run axis)
template <typename T> __device__ __forceinline__ int run_table_index(const AxisDesc<T>& A, int p) {
  return padded_to_table(A, p);
}

template <typename T, int ND, bool SK>
__global__ void __launch_bounds__(kTileThreads)
eval_tile_kernel(const EvalArgs<T, ND> a) {
  using Cfg = TileCfg<T, ND>;
  extern __shared__ __align__(16) char smem[];
  __shared__ AxisConst<T> consts[ND];
  __shared__ ipx_params params_s;

  // ---- per-CTA staging: constants, traced operands, knots and reciprocal widths ----
  if (threadIdx.x < ND) axis_consts(a.ax[threadIdx.x], consts[threadIdx.x]);
  if (threadIdx.x == 32) params_s = a.dp ? *a.dp : a.hp;
  AxisSm<T> S[ND];
  char* tiles = smem;
  if (SK) {
    T* base = reinterpret_cast<T*>(smem);
    int off = 0;
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      const int nk = a.ax[ax].nk;
      T* ks = base + off;
      T* rs = base + off + nk;
      for (int i = threadIdx.x; i < nk; i += kTileThreads) {
        T xi = a.ax[ax].knots[i];
        ks[i] = xi;
        rs[i] = i > 0 ? safe_recip(xi - a.ax[ax].knots[i - 1]) : T(0);
      }
      S[ax].knots = ks;
      S[ax].rdx = rs;
      off += 2 * nk;
    }
    tiles = smem + ((off * (int)sizeof(T) + 15) & ~15);
  } else {
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) { S[ax].knots = a.ax[ax].knots; S[ax].rdx = nullptr; }
  }
  __syncthreads();
  const ipx_params& P = params_s;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  char* tile = tiles + warp * Cfg::WARP_TILE_BYTES;
  const T* __restrict__ tab = a.tab[0];
  const AxisDesc<T>& RA = a.ax[ND - 1];  // run axis

  const int64_t step = (int64_t)gridDim.x * kTileThreads;
  int64_t qi = (int64_t)blockIdx.x * kTileThreads + threadIdx.x;

  // software prefetch of the query stream: the loads for the next round are in flight while
  // this round gathers and contracts
  T qn[ND];
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) qn[ax] = qi < a.nq ? a.ax[ax].q[qi] : T(0);

  for (; __any_sync(0xffffffffu, qi < a.nq); qi += step) {
    const bool active = qi < a.nq;
    T qc[ND];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) qc[ax] = qn[ax];
    {
      const int64_t qnext = qi + step;
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) qn[ax] = qnext < a.nq ? a.ax[ax].q[qnext] : T(0);
    }

    AxisLoc<T> L[ND];
    T w[ND][4];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      T dxi;
      locate_s(a.ax[ax], S[ax], consts[ax], P.axis[ax], qc[ax], L[ax], dxi);
      hermite_weights_r(L[ax], P.axis[ax].derivative, dxi, w[ax]);
    }

    // ---- group lanes by cell column (all axes but the last) ----
    unsigned long long key = 0;
#pragma unroll
    for (int ax = 0; ax < ND - 1; ++ax) key = key * (unsigned long long)(a.ax[ax].nk + 1) + (unsigned)L[ax].i;
    if (!active) key = ~0ull;
    const unsigned same = __match_any_sync(0xffffffffu, key);
    const bool is_leader = (__ffs(same) - 1) == lane;
    const unsigned leaders = __ballot_sync(0xffffffffu, is_leader && active);
    const int pz = L[ND - 1].i - 1;  // lower knot position on the run axis

    T res = T(0);
    if (__popc(leaders) <= kMaxGroups) {
      unsigned todo = __ballot_sync(0xffffffffu, active);
      while (todo) {
        const int ld = __ffs(todo) - 1;
        const unsigned grp = __shfl_sync(0xffffffffu, same, ld) & todo;
        const bool in_grp = (grp >> lane) & 1u;
        const int zmin = __reduce_min_sync(0xffffffffu, in_grp ? pz : 0x7fffffff);
        const bool sub = in_grp && (pz - zmin) <= kRunCap - 2;
        const unsigned subm = __ballot_sync(0xffffffffu, sub);
        const int zmax = __reduce_max_sync(0xffffffffu, sub ? pz : -0x7fffffff);
        const int extent = zmax - zmin + 2;  // nodes per column, <= kRunCap

        // column base node indices from the leader's corners
        int64_t colbase[Cfg::NC];
#pragma unroll
        for (int c = 0; c < Cfg::NC; ++c) {
          int64_t idx = 0;
#pragma unroll
          for (int ax = 0; ax < ND - 1; ++ax) {
            const int t0 = __shfl_sync(0xffffffffu, L[ax].t0, ld);
            const int t1 = __shfl_sync(0xffffffffu, L[ax].t1, ld);
            idx = idx * a.ax[ax].n + (((c >> ax) & 1) ? t1 : t0);
          }
          colbase[c] = idx * RA.n;
        }
        // ---- cooperative copy global -> shared: NC columns x extent nodes x CH chunks ----
        const int per_col = extent * Cfg::CH;
#pragma unroll
        for (int c = 0; c < Cfg::NC; ++c) {
          for (int r = lane; r < per_col; r += 32) {
            const int node = r / Cfg::CH;
            const int part = r - node * Cfg::CH;
            const int zt = run_table_index(RA, zmin + node);
            const float4 v = __ldg(reinterpret_cast<const float4*>(
                reinterpret_cast<const char*>(tab + (colbase[c] + zt) * Cfg::NT) + 16 * part));
            *reinterpret_cast<float4*>(tile + (c * kRunCap + node) * Cfg::STRIDE + 16 * part) = v;
          }
        }
        __syncwarp();
        if (sub) {
          const int nz0 = pz - zmin;
          res = contract_cubic<T, ND>(w, [&](int bits, T* r) {
            const int c = bits & (Cfg::NC - 1);
            const int node = nz0 + (bits >> (ND - 1));
            lds_record<T, Cfg::NT>(tile + (c * kRunCap + node) * Cfg::STRIDE, r);
          });
        }
        __syncwarp();
        todo &= ~subm;
      }
    } else if (active) {
      // ---- direct per-lane gather (incoherent queries) ----
      res = contract_cubic<T, ND>(w, [&](int bits, T* r) {
        load_record<T, ND, IPX_AOS>(a, node_index<T, ND>(a, L, bits), r);
      });
    }

    if (active) {
      a.out[qi] = apply_extrap<T, ND>(a, P, L, res);
      if (a.idx_out != nullptr) {
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) a.idx_out[(int64_t)ax * a.nq + qi] = L[ax].i;
      }
    }
  }
}

// dynamic shared memory the tile kernel needs
template <typename T, int ND> static size_t tile_smem_bytes(const EvalArgs<T, ND>& a, bool sk) {
  size_t knots = 0;
  if (sk) {
    for (int ax = 0; ax < ND; ++ax) knots += 2 * (size_t)a.ax[ax].nk * sizeof(T);
    knots = (knots + 15) & ~(size_t)15;
  }
  return knots + (size_t)kTileWarps * TileCfg<T, ND>::WARP_TILE_BYTES;
}

template <typename T, int ND, bool SK>
static cudaError_t launch_tile_sk(const EvalArgs<T, ND>& a, cudaStream_t s) {
  const size_t smem = tile_smem_bytes<T, ND>(a, SK);
  auto kern = eval_tile_kernel<T, ND, SK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t need = (a.nq + kTileThreads - 1) / kTileThreads;
  int64_t grid = (int64_t)sms * per_sm;  // persistent: one resident wave
  if (need < grid) grid = need;
  kern<<<(int)grid, kTileThreads, smem, s>>>(a);
  return cudaSuccess;
}

// true if (T, ND) has a tile kernel: records of at least 16 bytes
template <typename T, int ND> constexpr bool tile_supported() {
  return (1 << ND) * (int)sizeof(T) >= 16;
}

template <typename T, int ND>
static bool launch_tile(const EvalArgs<T, ND>& a, cudaStream_t s) {
  if constexpr (!tile_supported<T, ND>()) {
    return false;
  } else {
    int total_knots = 0;
    for (int ax = 0; ax < ND; ++ax) total_knots += a.ax[ax].nk;
    cudaError_t e = total_knots <= kSmemKnotCap ? launch_tile_sk<T, ND, true>(a, s)
                                                : launch_tile_sk<T, ND, false>(a, s);
    return e == cudaSuccess;
  }
}

}  // namespace ipx
