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

// cubic Hermite weights from the local coordinate pieces (same formulas as hermite_weights)
template <typename T>
__device__ __forceinline__ void hermite_weights_r(T q, T x0, T x1, int order, T dxi, T w[4]) {
  T dx = x1 - x0;
  T t = (q - x0) * dxi;
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
      res = fma(r[1], w[0][2 + ii], fma(r[0], w[0][ii], res));
    }
  } else if (ND == 2) {
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      T p[2] = {T(0), T(0)};
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        T r[4];
        load(ii | (jj << 1), r);
        p[0] = fma(r[1], w[1][2 + jj], fma(r[0], w[1][jj], p[0]));
        p[1] = fma(r[3], w[1][2 + jj], fma(r[2], w[1][jj], p[1]));
      }
      res = fma(p[1], w[0][2 + ii], fma(p[0], w[0][ii], res));
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
          for (int m = 0; m < 4; ++m)
            p[m] = fma(r[2 * m + 1], w[2][2 + kk], fma(r[2 * m], w[2][kk], p[m]));
        }
        v[0] = fma(p[1], w[1][2 + jj], fma(p[0], w[1][jj], v[0]));
        v[1] = fma(p[3], w[1][2 + jj], fma(p[2], w[1][jj], v[1]));
      }
      res = fma(v[1], w[0][2 + ii], fma(v[0], w[0][ii], res));
    }
  }
  return res;
}

// table index of knot position p along the LAST axis (run axis)
template <typename T> __device__ __forceinline__ int run_table_index(const AxisDesc<T>& A, int p) {
  return padded_to_table(A, p);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// knot-derived constants of one axis (need device data, so they live in shared memory);
// everything else comes straight from the kernel parameters (constant bank: no LSU traffic)
template <typename T> struct TileAxis {
  T x_first, inv_h, x_last, pad;
};

template <typename T, int ND, bool DEVP>
__global__ void __launch_bounds__(kTileThreads, 3)
eval_tile_kernel(const EvalArgs<T, ND> a) {
  using Cfg = TileCfg<T, ND>;
  constexpr int LPC = 32 / Cfg::NC;          // lanes that copy one column
  constexpr int NPR = LPC / Cfg::CH;         // nodes of a column copied per round
  static_assert(LPC % Cfg::CH == 0 && NPR >= 1, "column copy layout");
  extern __shared__ __align__(16) char smem[];
  __shared__ TileAxis<T> AX[ND];
  T* const smT = reinterpret_cast<T*>(smem);

  // ---- per-CTA staging: axis constants, traced operands, knots and reciprocal widths ----
  __shared__ ipx_params params_s;  // only used when the traced operands live in device memory
  if (threadIdx.x < ND) {
    const int ax = threadIdx.x;
    AxisConst<T> c;
    axis_consts(a.ax[ax], c);
    AX[ax].x_first = c.x_first; AX[ax].inv_h = c.inv_h; AX[ax].x_last = c.x_last;
  }
  if (DEVP && threadIdx.x == 32) params_s = *a.dp;
  const ipx_params& P = DEVP ? params_s : a.hp;
  int knot_elems = 0;
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    const int nk = a.ax[ax].nk;
    for (int i = threadIdx.x; i < nk; i += kTileThreads) {
      T xi = a.ax[ax].knots[i];
      smT[knot_elems + i] = xi;
      smT[knot_elems + nk + i] = i > 0 ? safe_recip(xi - a.ax[ax].knots[i - 1]) : T(0);
    }
    knot_elems += 2 * nk;
  }
  char* const tiles = smem + ((knot_elems * (int)sizeof(T) + 15) & ~15);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  char* const tile = tiles + warp * Cfg::WARP_TILE_BYTES;
  const char* __restrict__ tabc = reinterpret_cast<const char*>(a.tab[0]);
  const AxisDesc<T>& RA = a.ax[ND - 1];  // run axis
  // the run of nodes is contiguous in the table unless the run axis is periodic with a
  // permutation (un-sorted knots) or the run crosses the periodic seam
  const bool run_identity = !RA.periodic || RA.perm == nullptr;
  const bool want_idx = a.idx_out != nullptr;

  // this lane's role in the cooperative copy: column `ccol`, chunk `lane % LPC` of each round
  const int ccol = lane / LPC;
  const int cl = lane - ccol * LPC;
  const int cnode0 = cl / Cfg::CH;
  const int cpart = cl - cnode0 * Cfg::CH;
  const unsigned cdst0 = (unsigned)__cvta_generic_to_shared(tile) +
                         (ccol * kRunCap + cnode0) * Cfg::STRIDE + 16 * cpart;

  const int64_t step = (int64_t)gridDim.x * kTileThreads;
  const int64_t q0 = (int64_t)blockIdx.x * kTileThreads + (threadIdx.x & ~31);  // warp's first query
  if (q0 >= a.nq) return;  // whole warp idle (after the CTA barrier above)
  const int rounds = (int)((a.nq - q0 + step - 1) / step);

  // software prefetch of the query stream: the loads for the next round are in flight while
  // this round gathers and contracts
  int64_t qi = q0 + lane;
  T qn[ND];
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) qn[ax] = qi < a.nq ? a.ax[ax].q[qi] : T(0);

  for (int round = 0; round < rounds; ++round, qi += step) {
    const bool active = qi < a.nq;
    T qc[ND];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) qc[ax] = qn[ax];
    {
      const int64_t qnext = qi + step;
      const bool more = qnext < a.nq;
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) qn[ax] = more ? a.ax[ax].q[qnext] : T(0);
    }

    // ---- locate: wrap, interval lookup in shared memory, corner indices ----
    int li[ND], t0[ND], t1[ND];
    T x0[ND], x1[ND], dxi[ND];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      const AxisDesc<T>& A = a.ax[ax];
      T q = qc[ax];
      if (A.wrap) {
        const T Pd = T(fabs(P.axis[ax].period));
        T r = q;
        r = q < T(0) ? q + Pd : r;      // [-P,0): fmod leaves q, the sign fix adds P
        r = q >= Pd ? q - Pd : r;       // [P,2P]: exact (Sterbenz)
        if (!(q >= -Pd && q <= Pd + Pd)) r = py_mod(q, Pd);  // anything else: the exact remainder
        q = r;
        qc[ax] = q;
      }
      const int nk = A.nk;
      int koff = 0;
#pragma unroll
      for (int k = 0; k < ax; ++k) koff += 2 * a.ax[k].nk;
      const T* ks = smT + koff;
      const T xf = AX[ax].x_first, ih = AX[ax].inv_h;
      int i = min(max(t_floor_to_int((q - xf) * ih), 0), nk - 2) + 1;
      T lo = ks[i - 1], hi = ks[i];
      // exact check of the guess against the knots (NaN fails it); the rare miss re-searches
      if (!((q >= lo || i == 1) && (q < hi || i == nk - 1))) {
        i = bin_index(ks, nk, q, xf, ih, lo, hi);
      }
      li[ax] = i; x0[ax] = lo; x1[ax] = hi;
      dxi[ax] = ks[nk + i];
      if (A.periodic) {
        int s0 = i - 2; s0 += s0 < 0 ? A.n : 0;
        int s1 = i - 1; s1 -= s1 >= A.n ? A.n : 0;
        const int32_t* pm = A.perm;
        t0[ax] = pm ? pm[s0] : s0;
        t1[ax] = pm ? pm[s1] : s1;
      } else {
        t0[ax] = i - 1; t1[ax] = i;
      }
    }

    // ---- group lanes by cell column (all axes but the last) ----
    unsigned key = 0;
    bool key_ok = true;  // 32-bit keys suffice unless the knot vectors are huge
#pragma unroll
    for (int ax = 0; ax < ND - 1; ++ax) {
      const unsigned m = (unsigned)a.ax[ax].nk + 1u;
      key_ok = key_ok && ((unsigned long long)key * m < 0xffff0000ull);
      key = key * m + (unsigned)li[ax];
    }
    if (!active) key = 0xffffffffu;
    const unsigned same = __match_any_sync(0xffffffffu, key);
    const unsigned leaders = __ballot_sync(0xffffffffu, ((__ffs(same) - 1) == lane) && active);
    const int pz = li[ND - 1] - 1;  // lower knot position on the run axis
    const bool coherent = key_ok && __popc(leaders) <= kMaxGroups;

    T w[ND][4];
    T res = T(0);
    if (coherent) {
      unsigned todo = __ballot_sync(0xffffffffu, active);
      bool have_w = false;
      while (todo) {
        const int ld = __ffs(todo) - 1;
        const unsigned grp = __shfl_sync(0xffffffffu, same, ld) & todo;
        const bool in_grp = (grp >> lane) & 1u;
        const int zmin = __reduce_min_sync(0xffffffffu, in_grp ? pz : 0x7fffffff);
        const bool sub = in_grp && (pz - zmin) <= kRunCap - 2;
        const unsigned subm = __ballot_sync(0xffffffffu, sub);
        const int zmax = __reduce_max_sync(0xffffffffu, sub ? pz : -0x7fffffff);
        const int extent = zmax - zmin + 2;  // nodes per column, <= kRunCap

        // node index of (this lane's copy column, run position 0), from the leader's corners
        int colnode = 0;
#pragma unroll
        for (int ax = 0; ax < ND - 1; ++ax) {
          const int c0 = __shfl_sync(0xffffffffu, t0[ax], ld);
          const int c1 = __shfl_sync(0xffffffffu, t1[ax], ld);
          colnode = colnode * a.ax[ax].n + (((ccol >> ax) & 1) ? c1 : c0);
        }
        // table position of the first node of the run, if the run is contiguous in the table
        int zt0 = zmin;
        bool contiguous = run_identity;
        if (RA.periodic) {
          zt0 = zmin - 1;
          contiguous = contiguous && zmin >= 1 && (zmin + extent - 1) <= RA.n;
        }
        // ---- cooperative copy global -> shared, asynchronous (cp.async, no register staging)
        if (contiguous) {
          const char* src = tabc + ((size_t)colnode * RA.n + zt0) * Cfg::REC_BYTES + 16 * cl;
          unsigned dst = cdst0;
          for (int node = cnode0; node < extent; node += NPR) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src) : "memory");
            src += NPR * Cfg::REC_BYTES;
            dst += NPR * Cfg::STRIDE;
          }
        } else {
          for (int node = cnode0; node < extent; node += NPR) {
            const int zt = run_table_index(RA, zmin + node);
            cp_async16(tile + (ccol * kRunCap + node) * Cfg::STRIDE + 16 * cpart,
                       tabc + ((size_t)colnode * RA.n + zt) * Cfg::REC_BYTES + 16 * cpart);
          }
        }
        // the Hermite weights do not depend on the table: compute them under the copy
        if (!have_w) {
#pragma unroll
          for (int ax = 0; ax < ND; ++ax)
            hermite_weights_r(qc[ax], x0[ax], x1[ax], P.axis[ax].derivative, dxi[ax], w[ax]);
          have_w = true;
        }
        cp_async_wait_all();
        __syncwarp();
        if (sub) {
          const char* rec0 = tile + (pz - zmin) * Cfg::STRIDE;
          res = contract_cubic<T, ND>(w, [&](int bits, T* r) {
            const int c = bits & (Cfg::NC - 1);
            const int kk = bits >> (ND - 1);
            lds_record<T, Cfg::NT>(rec0 + (c * kRunCap + kk) * Cfg::STRIDE, r);
          });
        }
        __syncwarp();
        todo &= ~subm;
      }
    } else if (active) {
      // ---- direct per-lane gather (incoherent queries) ----
#pragma unroll
      for (int ax = 0; ax < ND; ++ax)
        hermite_weights_r(qc[ax], x0[ax], x1[ax], P.axis[ax].derivative, dxi[ax], w[ax]);
      res = contract_cubic<T, ND>(w, [&](int bits, T* r) {
        size_t node = (bits & 1) ? t1[0] : t0[0];
#pragma unroll
        for (int ax = 1; ax < ND; ++ax) node = node * a.ax[ax].n + (((bits >> ax) & 1) ? t1[ax] : t0[ax]);
        const T* p = reinterpret_cast<const T*>(tabc) + node * Cfg::NT;
        if (ND == 1) ldg_rec2(p, r);
        if (ND == 2) ldg_rec4(p, r);
        if (ND == 3) ldg_rec8(p, r);
      });
    }

    if (active) {
      // extrapolation fills, x then y then z, later axes win (_spline.py:1101-1103)
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) {
        if (a.ax[ax].wrap) continue;
        const ipx_axis_params& pa = P.axis[ax];
        if (!pa.keep_lo && qc[ax] < AX[ax].x_first) res = T(pa.fill_lo);
        if (!pa.keep_hi && qc[ax] > AX[ax].x_last) res = T(pa.fill_hi);
      }
      a.out[qi] = res;
      if (want_idx) {
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) a.idx_out[(int64_t)ax * a.nq + qi] = li[ax];
      }
    }
  }
}

// dynamic shared memory the tile kernel needs
template <typename T, int ND> static size_t tile_smem_bytes(const EvalArgs<T, ND>& a) {
  size_t knots = 0;
  for (int ax = 0; ax < ND; ++ax) knots += 2 * (size_t)a.ax[ax].nk * sizeof(T);
  knots = (knots + 15) & ~(size_t)15;
  return knots + (size_t)kTileWarps * TileCfg<T, ND>::WARP_TILE_BYTES;
}

// true if (T, ND) has a tile kernel: records of at least 16 bytes
template <typename T, int ND> constexpr bool tile_supported() {
  return (1 << ND) * (int)sizeof(T) >= 16;
}

// Launches the tile kernel if it applies (record >= 16 B, knot vectors fit in shared memory);
// returns false if the caller should use the generic kernel instead.
template <typename T, int ND>
static bool launch_tile(const EvalArgs<T, ND>& a, cudaStream_t s) {
  if constexpr (!tile_supported<T, ND>()) {
    return false;
  } else {
    int total_knots = 0;
    for (int ax = 0; ax < ND; ++ax) total_knots += a.ax[ax].nk;
    if (total_knots > kSmemKnotCap) return false;
    const size_t smem = tile_smem_bytes<T, ND>(a);
    auto kern = a.dp ? eval_tile_kernel<T, ND, true> : eval_tile_kernel<T, ND, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return false;
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileThreads, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t need = (a.nq + kTileThreads - 1) / kTileThreads;
    int64_t grid = (int64_t)sms * per_sm;  // persistent: one resident wave
    if (need < grid) grid = need;
    kern<<<(int)grid, kTileThreads, smem, s>>>(a);
    return true;
  }
}

}  // namespace ipx
