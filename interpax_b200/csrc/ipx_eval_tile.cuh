// ipx_eval_tile.cuh -- warp-cooperative, shared-memory-staged evaluation of the cubic family
// on AoS-packed tables (the Interpolator1D/2D/3D path).  Included by ipx_eval.cu.
//
// Why.  A cubic query needs 2^d corner records (8 x 64 B in 3-D f64 = 512 B of operands).
// Delivering operands to registers costs one LSU wavefront per 128 B per warp no matter where
// they come from (L1 or shared memory), so a thread-per-query gather is bound by the LSU/L1
// pipe -- 4 cycles per query per SM in 3-D f64 -- long before HBM
// (profiles/ncu_eval3d_r01_baseline.txt).  Spatially coherent (cell-sorted) queries put
// neighbouring queries into the same or the next cell along the fastest table axis.  So:
//   1. a warp takes 64 consecutive queries, two per lane;
//   2. lanes are grouped by cell "column" (interval indices on all axes but the last);
//   3. the group's run of nodes -- 2^(d-1) columns x (run+1) nodes, contiguous along the last
//      axis -- is copied global -> shared with cp.async, full-line 128-bit accesses (each byte
//      crosses L1 once per warp, not once per lane);
//   4. every lane walks the run one node LAYER at a time (a layer = the 2^(d-1) records at one
//      position of the last axis), reads each record from shared memory ONCE and applies it
//      to both of its queries.  Two queries in the same cell share both layers, in adjacent
//      cells one, so operand traffic per query drops by up to 2x;
//   5. the contraction is separable and runs last-axis-last: per layer, contract the other
//      axes into (value part, last-axis-slope part), then combine with the last-axis weights.
// A second query that is not in its partner's cell or the one above it, and warps whose
// queries are spread over many columns (incoherent order), use the same layer arithmetic on
// the lane's own query with a direct gather, so a query's result never depends on which path
// served it (tests/test_gpu_parity.py::test_full_size_properties_c4 checks bitwise invariance
// under permutation of the queries).
//
// Knot vectors and reciprocal cell widths are staged in shared memory once per CTA (the grid
// is persistent): the interval lookup costs no global loads and the local coordinate no
// division per query.  1/dx is produced by the same IEEE division as in the generic kernel.
#pragma once
#include <atomic>
#include <cstdlib>

namespace ipx {

// One CTA per SM: the knot vectors (and their reciprocal widths) are staged once per SM, not once
// per resident CTA, and what shared memory is not needed stays L1 -- the direct gathers of
// incoherent (random-order) streams live on L1 capacity.  The register file holds 16 warps of a
// float64 kernel (128 registers) or 20 of a float32 one (96; measured against 16 and 24 warps,
// profiles/ in the DESIGN table); the launch uses fewer warps when long knot vectors leave less
// room for the per-warp tiles.
#ifndef IPX_TILE_PAD
#define IPX_TILE_PAD 1
#endif
#ifndef IPX_TILE_F64_THREADS
#define IPX_TILE_F64_THREADS 512
#endif
// smallest node spacing for which the 8 / ch lane groups of a quarter warp (ch 16-byte chunks
// each, node stride = stride_chunks chunks) write 8 different bank groups
__host__ __device__ constexpr int copy_node_spacing(int stride_chunks, int ch) {
  for (int sp = 1; sp <= 8; ++sp) {
    unsigned used = 0;
    bool ok = true;
    for (int g = 0; g < 8 / ch; ++g)
      for (int c = 0; c < ch; ++c) {
        const unsigned bit = 1u << ((g * sp * stride_chunks + c) % 8);
        ok = ok && (used & bit) == 0;
        used |= bit;
      }
    if (ok) return sp;
  }
  return 1;
}
template <typename T> constexpr int tile_max_threads() { return sizeof(T) == 4 ? 640 : IPX_TILE_F64_THREADS; }
constexpr int kMaxGroups = 4;       // more distinct columns than this in a warp -> direct path
constexpr int kSmemKnotCap = 6144;  // total knots (all axes) staged in shared memory

// RC = nodes of one run held per column (16: dense queries, several per cell; 32: sparse ones)
template <typename T, int ND, int RC = 16, int QPL = 2> struct TileCfg {
  static constexpr int NT = 1 << ND;                       // values per record
  static constexpr int NC = 1 << (ND - 1);                 // columns per group
  static constexpr int NCA = ND > 1 ? ND - 1 : 1;          // column axes (array extent)
  static constexpr int REC_BYTES = NT * (int)sizeof(T);
  static constexpr int CH = REC_BYTES / 16;                // 16-byte chunks per record
  static constexpr int STRIDE = REC_BYTES + (IPX_TILE_PAD && REC_BYTES >= 32 ? 16 : 0);  // bank-conflict padding
  static constexpr int RUN_BYTES = NC * RC * STRIDE;         // node runs of one group
  static constexpr int QSLOT = QPL * (int)sizeof(T);         // a lane's coordinates on one axis
  static constexpr int QAXIS = 32 * QSLOT;
  static constexpr int WARP_TILE_BYTES = RUN_BYTES + ((ND * QAXIS + 15) & ~15);  // + next round's queries
};

// table index of knot position p along the LAST axis (run axis)
template <typename T, int WM, int AX>
__device__ __forceinline__ int run_table_index(const AxisDesc<T>& A, int p) {
  if constexpr (WM >= 0) {
    if constexpr (((WM >> AX) & 1) == 0) return p;
    int s = p - 1;
    if (s < 0) s += A.n;
    if (s >= A.n) s -= A.n;
    return s;
  } else {
    return padded_to_table(A, p);
  }
}

// knot-derived constants of one axis (need device data, so they live in shared memory);
// everything else comes straight from the kernel parameters (constant bank: no LSU traffic)
template <typename T> struct TileAxis {
  T x_first, inv_h, x_last, pad;
};

// one located query: weights formed, indices kept
template <typename T, int ND> struct QState {
  T w[ND][4];
  int li[ND];                      // interval index per axis
  int t0[TileCfg<T, ND>::NCA];     // column-axis corner table indices
  int t1[TileCfg<T, ND>::NCA];
  unsigned fill;                   // bit 2ax: q < knots[0], bit 2ax+1: q > knots[-1] (axes that fill)
};

__device__ __forceinline__ void cp_async16_s(unsigned smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
// query coordinates are read once: evict-first in L2, so the stream does not push table lines out
// (the register path used __ldcs for the same reason)
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
template <int BYTES>
__device__ __forceinline__ void cp_async_s(unsigned smem_dst, const void* gmem_src, unsigned long long pol) {
  if constexpr (BYTES == 16) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(smem_dst), "l"(gmem_src), "l"(pol) : "memory");
  } else if constexpr (BYTES == 8) {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;\n" ::"r"(smem_dst), "l"(gmem_src), "l"(pol) : "memory");
  } else {
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 4, %2;\n" ::"r"(smem_dst), "l"(gmem_src), "l"(pol) : "memory");
  }
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// DEVP: traced operands are read from device memory.  ORD0: every derivative order is 0 (known
// on the host when the operands are host-resident), which drops the order switch from the loop.
// TWO: two consecutive queries per lane (dense streams, several queries per cell: the pair
// shares node layers); otherwise one per lane (sparse streams, about one query per cell or
// fewer: a warp's queries then span fewer cells, fit one run and keep every lane busy).
// WM: -1, or the set of periodic axes as a compile-time mask (bit ax).  The mask form is used
// when every periodic axis wraps its queries and maps padded positions to the table without a
// permutation (sorted knots, un-padded tables: what Interpolator1D/2D/3D passes); the wrap,
// seam and fill code of the other axes then drops out of the loop instead of being branched
// around per query.
// PM: with WM >= 0, the axes whose padded knot positions map back into an un-padded table (what
// the Interpolator classes pass: PM == WM); 0 when the periodic tables arrive padded (the
// functional form, which differentiates the padded table: _spline.py:924-938, 1027-1040).
template <typename T, int ND, bool DEVP, bool ORD0, int RC, bool TWO = true, int WM = -1, int PM = WM>
__global__ void __launch_bounds__(tile_max_threads<T>(), 1)
eval_tile_kernel(const EvalArgs<T, ND> a) {
  constexpr int QPL = TWO ? 2 : 1;  // queries per lane per round
  using Cfg = TileCfg<T, ND, RC, QPL>;
  constexpr bool CTM = WM >= 0;
#define IPXT_WRAP(AXI) (CTM ? (((WM >> (AXI)) & 1) != 0) : (a.ax[AXI].wrap != 0))
#define IPXT_PERIODIC(AXI) (CTM ? (((PM >> (AXI)) & 1) != 0) : (a.ax[AXI].periodic != 0))
#define IPXT_PERM(AXI) (CTM ? (const int32_t*)nullptr : a.ax[AXI].perm)
  constexpr int kRunCap = RC;
  constexpr int LPC = 32 / Cfg::NC;          // lanes that copy one column
  constexpr int NPR = LPC / Cfg::CH;         // nodes of a column copied per round
  static_assert(LPC % Cfg::CH == 0 && NPR >= 1, "column copy layout");
  // One cp.async instruction writes, per column, the 16-byte chunks of NPR nodes; a quarter
  // warp's 8 chunks land in one shared-memory wavefront only if they hit 8 different 16-byte bank
  // groups.  With the padded node stride, nodes that are neighbours collide (80 B: node n+1
  // wraps onto node n's first group), nodes SP apart do not, SP = the spacing for which
  // SP * STRIDE advances the bank group by exactly one record: lane group g copies nodes
  // base + j + SP * g, j = 0 .. SP-1.
  constexpr int SP = copy_node_spacing(Cfg::STRIDE / 16, Cfg::CH);
  extern __shared__ __align__(16) char smem[];
  __shared__ TileAxis<T> AX[ND];
  __shared__ ipx_params params_s;  // only used when the traced operands live in device memory
  T* const smT = reinterpret_cast<T*>(smem);
  if (skip_launch(a)) return;

  // ---- per-CTA staging: axis constants, traced operands, knots and reciprocal widths ----
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
    for (int i = threadIdx.x; i < nk; i += (int)blockDim.x) {
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
  const bool run_identity = !IPXT_PERIODIC(ND - 1) || IPXT_PERM(ND - 1) == nullptr;
  const bool want_idx = a.idx_out != nullptr;

  // this lane's role in the cooperative copy: column `ccol`, chunk `lane % LPC` of each round
  const int ccol = lane / LPC;
  const int cl = lane - ccol * LPC;
  const int cnode0 = cl / Cfg::CH;
  const int cpart = cl - cnode0 * Cfg::CH;
  const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tile);
  const unsigned cdst0 = tile_s + (ccol * kRunCap + SP * cnode0) * Cfg::STRIDE + 16 * cpart;

  // ---- locate both queries of the lane ---------------------------------------------------
  // Written as three straight-line phases so that the 2 x ND independent dependency chains
  // (wrap -> guess -> knot loads -> check -> weights) interleave instead of running one after
  // the other: (A) wrap, guess and knot loads for every (query, axis) with the exceptional
  // cases only FLAGGED; (B) one branch that repairs flagged entries (rare: |q| beyond one
  // period, a guess off by a cell, NaN); (C) weights, fill flags and corner indices.
  // The next round's coordinates wait in shared memory, not in registers (see the loop below),
  // which is what lets the 2 x 3 chains of a 3-D pair fit the 128-register budget.
  auto locate_range = [&](const T (&qraw)[2][ND], T (&q)[2][ND], T (&lo)[2][ND], T (&hi)[2][ND], T (&dxi)[2][ND],
                          QState<T, ND> (&S)[2], int (&pz)[2], const int k_lo, const int k_hi) {
    int ii[2][ND];
    unsigned bad = 0;  // bit 2*ax + k set: entry (k, ax) needs the slow path
    // (A)
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      const AxisDesc<T>& A = a.ax[ax];
      const int nk = A.nk;
      int koff = 0;
#pragma unroll
      for (int k = 0; k < ax; ++k) koff += 2 * a.ax[k].nk;
      const T* ks = smT + koff;
      const T xf = AX[ax].x_first, ih = AX[ax].inv_h;
      const T Pd = T(fabs(P.axis[ax].period));
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k < k_lo || k >= k_hi) continue;
        T v = qraw[k][ax];
        bool wrap_bad = false;
        if (IPXT_WRAP(ax)) v = wrap_one_period(v, Pd, wrap_bad);
        const int i = min(max(t_floor_to_int((v - xf) * ih), 0), nk - 2) + 1;
        const T l = ks[i - 1], h = ks[i];
        // exact check of the guess against the knots; NaN fails it
        const bool ok = (v >= l || i == 1) && (v < h || i == nk - 1);
        if (wrap_bad || !ok) bad |= 1u << (2 * ax + k);
        q[k][ax] = v; lo[k][ax] = l; hi[k][ax] = h; ii[k][ax] = i;
        dxi[k][ax] = ks[nk + i];
      }
    }
    // (B)
    if (bad) {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) {
        const AxisDesc<T>& A = a.ax[ax];
        const int nk = A.nk;
        int koff = 0;
#pragma unroll
        for (int k = 0; k < ax; ++k) koff += 2 * a.ax[k].nk;
        const T* ks = smT + koff;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k < k_lo || k >= k_hi) continue;
          if (bad & (1u << (2 * ax + k))) {
            T v = qraw[k][ax];
            if (IPXT_WRAP(ax)) v = py_mod_slow(v, T(fabs(P.axis[ax].period)));
            const int i = bin_index_slow(ks, nk, v, AX[ax].x_first, AX[ax].inv_h);
            q[k][ax] = v; ii[k][ax] = i;
            lo[k][ax] = ks[i - 1]; hi[k][ax] = ks[i]; dxi[k][ax] = ks[nk + i];
          }
        }
      }
    }
    // (C) corner indices; the weights follow in form_weights, after the run copy has been issued
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k < k_lo || k >= k_hi) continue;
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) {
        const AxisDesc<T>& A = a.ax[ax];
        const int i = ii[k][ax];
        S[k].li[ax] = i;
        if (ax < ND - 1) {
          int s0 = i - 2; s0 += s0 < 0 ? A.n : 0;
          int s1 = i - 1; s1 -= s1 >= A.n ? A.n : 0;
          S[k].t0[ax] = IPXT_PERIODIC(ax) ? s0 : i - 1;
          S[k].t1[ax] = IPXT_PERIODIC(ax) ? s1 : i;
        }
      }
      pz[k] = S[k].li[ND - 1] - 1;
    }
    // un-sorted periodic knots: sorted position -> table index (rare; identity is passed as NULL)
#pragma unroll
    for (int ax = 0; ax < ND - 1; ++ax) {
      const int32_t* pm = IPXT_PERIODIC(ax) ? IPXT_PERM(ax) : nullptr;
      if (pm != nullptr) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if (k < k_lo || k >= k_hi) continue;
          S[k].t0[ax] = pm[S[k].t0[ax]];
          S[k].t1[ax] = pm[S[k].t1[ax]];
        }
      }
    }
  };

  // column key = interval indices of the column axes in mixed radix; it must stay below the
  // "inactive lane" key 0xffffffff (checked once, the radices are uniform)
  bool key_ok = true;
  {
    unsigned long long span = 1;
#pragma unroll
    for (int ax = 0; ax < ND - 1; ++ax) {
      span *= (unsigned long long)a.ax[ax].nk + 1ull;
      key_ok = key_ok && span < 0xffff0000ull;
    }
  }
  // (D) weights and fill flags of located queries
  auto form_weights = [&](const T (&q)[2][ND], const T (&lo)[2][ND], const T (&hi)[2][ND], const T (&dxi)[2][ND],
                          QState<T, ND> (&S)[2], const int k_hi) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k >= k_hi) continue;
      S[k].fill = 0;
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) {
        const int i = S[k].li[ax];
        hermite_weights_r<T, ORD0>(q[k][ax], lo[k][ax], hi[k][ax], P.axis[ax].derivative, dxi[k][ax],
                                   S[k].w[ax]);
        const ipx_axis_params& pa = P.axis[ax];
        // a query outside the knots sits in an end cell: the compares run only there
        if (!IPXT_WRAP(ax) && (i == 1 || i == a.ax[ax].nk - 1)) {
          const bool flo = !pa.keep_lo && q[k][ax] < AX[ax].x_first;
          const bool fhi = !pa.keep_hi && q[k][ax] > AX[ax].x_last;
          S[k].fill |= (flo ? 1u : 0u) << (2 * ax);
          S[k].fill |= (fhi ? 2u : 0u) << (2 * ax);
        }
      }
    }
  };

  auto column_key = [&](const QState<T, ND>& S) {
    unsigned key = 0;
#pragma unroll
    for (int ax = 0; ax < ND - 1; ++ax) key = key * ((unsigned)a.ax[ax].nk + 1u) + (unsigned)S.li[ax];
    return key;
  };

  // combine a layer with the run-axis weights of corner `kk` (0 = lower node, 1 = upper)
  auto finish_layer = [&](const QState<T, ND>& S, const T (&A)[2], int kk, T acc) {
    return fma(A[0], S.w[ND - 1][kk], fma(A[1], S.w[ND - 1][2 + kk], acc));
  };

  // direct per-lane gather of one query (incoherent queries, second queries that cannot ride)
  auto eval_direct = [&](const QState<T, ND>& S, int pz) {
    T res = T(0);
    WeightRefs<T, ND, 1> W;
    W.w[0] = S.w;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const int zt = run_table_index<T, PM, ND - 1>(RA, pz + kk);
      T A[1][2];
      layer_contract<T, ND, 1>(
          [&](int c, T* r) {
            size_t node = 0;
#pragma unroll
            for (int ax = 0; ax < ND - 1; ++ax) node = node * a.ax[ax].n + (((c >> ax) & 1) ? S.t1[ax] : S.t0[ax]);
            node = node * RA.n + zt;
            ldg_record_wide<T, Cfg::NT>(reinterpret_cast<const T*>(tabc) + node * Cfg::NT, r);
          },
          W, A);
      res = finish_layer(S, A[0], kk, res);
    }
    return res;
  };

  // extrapolation fills, x then y then z, later axes win (_spline.py:1101-1103)
  auto filled = [&](const QState<T, ND>& S, T res) {
    if (S.fill) {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) {
        if (S.fill & (1u << (2 * ax))) res = T(P.axis[ax].fill_lo);
        if (S.fill & (2u << (2 * ax))) res = T(P.axis[ax].fill_hi);
      }
    }
    return res;
  };
  auto store_idx = [&](const QState<T, ND>& S, int64_t qi) {
    if (want_idx) {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) a.idx_out[(int64_t)ax * a.nq + qi] = S.li[ax];
    }
  };
  auto store_q = [&](const QState<T, ND>& S, int64_t qi, T res) {
    __stcs(a.out + qi, filled(S, res));
    store_idx(S, qi);
  };
  // a lane's two results are adjacent: one store of the pair (qa is even; needs an aligned `out`)
  const bool pair_store = TWO && (reinterpret_cast<uintptr_t>(a.out) & (2 * sizeof(T) - 1)) == 0;
  auto store_pair = [&](const QState<T, ND>& S0, const QState<T, ND>& S1, int64_t qi, T r0, T r1) {
    r0 = filled(S0, r0);
    r1 = filled(S1, r1);
    if constexpr (sizeof(T) == 8) {
      __stcs(reinterpret_cast<double2*>(a.out + qi), make_double2(r0, r1));
    } else {
      __stcs(reinterpret_cast<float2*>(a.out + qi), make_float2(r0, r1));
    }
    store_idx(S0, qi);
    store_idx(S1, qi + 1);
  };

  // ---- persistent loop: 64 consecutive queries per warp per round ------------------------
  const int64_t step = (int64_t)gridDim.x * blockDim.x * QPL;
  const int64_t w0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + warp) * (32 * QPL);  // warp's first query
  if (w0 >= a.nq) return;  // whole warp idle (after the CTA barrier above)
  const int rounds = (int)((a.nq - w0 + step - 1) / step);

  int64_t qa = w0 + QPL * lane;  // this lane's first query; its second (TWO) is qa + 1
  // The coordinates of the next round travel global -> shared with cp.async into the lane's own
  // slots (QSLOT bytes per axis) and are read back at the top of that round, so they cost no
  // registers while the current round runs.  Each lane reads its slots before it issues the
  // copy that overwrites them, and nobody else touches them: no warp synchronisation needed.
  char* const qslot = tile + Cfg::RUN_BYTES + Cfg::QSLOT * lane;
  const unsigned qslot_s = (unsigned)__cvta_generic_to_shared(qslot);
  auto prefetch = [&](int64_t q) {
    if (q + QPL - 1 < a.nq) {  // q is a multiple of QPL and the arrays are 16-byte aligned
      const unsigned long long pol = l2_evict_first_policy();
#pragma unroll
      for (int ax = 0; ax < ND; ++ax)
        cp_async_s<Cfg::QSLOT>(qslot_s + ax * Cfg::QAXIS, a.ax[ax].q + q, pol);
    } else {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) {
        T* d = reinterpret_cast<T*>(qslot + ax * Cfg::QAXIS);
        d[0] = q < a.nq ? a.ax[ax].q[q] : T(0);
        if constexpr (TWO) d[1] = T(0);
      }
    }
  };
  prefetch(qa);

  for (int round = 0; round < rounds; ++round, qa += step) {
    const bool act_a = qa < a.nq;
    const bool act_b = TWO && qa + 1 < a.nq;
    T qc[2][ND];
    cp_async_wait_all();
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      const T* v = reinterpret_cast<const T*>(qslot + ax * Cfg::QAXIS);
      if constexpr (TWO) {
        if constexpr (sizeof(T) == 8) {
          const double2 w = *reinterpret_cast<const double2*>(v);
          qc[0][ax] = T(w.x); qc[1][ax] = T(w.y);
        } else {
          const float2 w = *reinterpret_cast<const float2*>(v);
          qc[0][ax] = T(w.x); qc[1][ax] = T(w.y);
        }
      } else {
        qc[0][ax] = v[0]; qc[1][ax] = T(0);
      }
    }
    if (round + 1 < rounds) prefetch(qa + step);

    QState<T, ND> SS[2];
    int pzz[2];
    T lq[2][ND], llo[2][ND], lhi[2][ND], ldxi[2][ND];
    locate_range(qc, lq, llo, lhi, ldxi, SS, pzz, 0, QPL);
    if constexpr (!TWO) {
      pzz[1] = pzz[0];
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) SS[1].li[ax] = SS[0].li[ax];
    }
    QState<T, ND>& Sa = SS[0];
    QState<T, ND>& Sb = SS[1];
    const int pza = pzz[0], pzb = pzz[1];

    unsigned ka = column_key(Sa);
    const unsigned kb = column_key(Sb);
    // b rides with a if it is in a's cell or the next one up the run axis
    const int dz = pzb - pza;
    const bool ride = act_b && kb == ka && (dz == 0 || dz == 1);
    const bool same = dz == 0;
    if (!act_a) ka = 0xffffffffu;
    const unsigned mates = __match_any_sync(0xffffffffu, ka);
    const unsigned leaders = __ballot_sync(0xffffffffu, ((__ffs(mates) - 1) == lane) && act_a);
    const bool coherent = key_ok && __popc(leaders) <= kMaxGroups;
    const int ptop = ride ? pzb : pza;  // highest cell this lane touches in the group pass

    // One group pass = the lanes of one column whose cells fit one run.  begin_pass picks the
    // next group and ISSUES the copy of its node run (global -> shared, cp.async, no register
    // staging); the caller waits for it after doing something else.
    unsigned todo = 0, subm = 0;
    int zmin = 0;
    bool sub = false;
    auto begin_pass = [&]() {
      const int ld = __ffs(todo) - 1;
      const unsigned grp = __shfl_sync(0xffffffffu, mates, ld) & todo;
      const bool in_grp = (grp >> lane) & 1u;
      zmin = __reduce_min_sync(0xffffffffu, in_grp ? pza : 0x7fffffff);
      sub = in_grp && (ptop - zmin) <= kRunCap - 2;
      subm = __ballot_sync(0xffffffffu, sub);
      const int zmax = __reduce_max_sync(0xffffffffu, sub ? ptop : -0x7fffffff);
      const int extent = zmax - zmin + 2;  // nodes per column, <= kRunCap

      // node index of (this lane's copy column, run position 0), from the leader's corners
      int colnode = 0;
#pragma unroll
      for (int ax = 0; ax < ND - 1; ++ax) {
        const int c0 = __shfl_sync(0xffffffffu, Sa.t0[ax], ld);
        const int c1 = __shfl_sync(0xffffffffu, Sa.t1[ax], ld);
        colnode = colnode * a.ax[ax].n + (((ccol >> ax) & 1) ? c1 : c0);
      }
      // table position of the first node of the run, if the run is contiguous in the table
      int zt0 = zmin;
      bool contiguous = run_identity;
      if (IPXT_PERIODIC(ND - 1)) {
        zt0 = zmin - 1;
        contiguous = contiguous && zmin >= 1 && (zmin + extent - 1) <= RA.n;
      }
      if (contiguous) {
        const char* src =
            tabc + ((size_t)colnode * RA.n + zt0) * Cfg::REC_BYTES + (SP * cnode0) * Cfg::REC_BYTES + 16 * cpart;
        unsigned dst = cdst0;
        for (int node = SP * cnode0; node < extent + SP * cnode0; node += SP * NPR) {
#pragma unroll
          for (int j = 0; j < SP; ++j)
            if (node + j < extent) cp_async16_s(dst + j * Cfg::STRIDE, src + j * Cfg::REC_BYTES);
          src += SP * NPR * Cfg::REC_BYTES;
          dst += SP * NPR * Cfg::STRIDE;
        }
      } else {
        for (int node = cnode0; node < extent; node += NPR) {
          const int zt = run_table_index<T, PM, ND - 1>(RA, zmin + node);
          cp_async16_s(tile_s + (ccol * kRunCap + node) * Cfg::STRIDE + 16 * cpart,
                       tabc + ((size_t)colnode * RA.n + zt) * Cfg::REC_BYTES + 16 * cpart);
        }
      }
    };
    if (coherent) {
      todo = __ballot_sync(0xffffffffu, act_a);
      if (todo) begin_pass();
    }
    // the weights are formed while the first run copy is in flight
    form_weights(lq, llo, lhi, ldxi, SS, QPL);
    if constexpr (!TWO) SS[1] = SS[0];

    T res_a = T(0), res_b = T(0);
    bool done_b = !act_b;
    if (coherent) {
      while (todo) {
        cp_async_wait_all();
        __syncwarp();
        if (sub) {
          const char* lay = tile + (pza - zmin) * Cfg::STRIDE;  // layer of a's lower node
          auto tile_loader = [](const char* base) {
            return [base](int c, T* r) { lds_record<T, Cfg::NT>(base + c * (kRunCap * Cfg::STRIDE), r); };
          };
          if (ride) {
            WeightRefs<T, ND, 2> W2;
            W2.w[0] = Sa.w;
            W2.w[1] = Sb.w;
            T A[2][2];
            // layer 0: lower node of a's cell (and of b's, if b shares the cell)
            layer_contract<T, ND, 2>(tile_loader(lay), W2, A);
            res_a = finish_layer(Sa, A[0], 0, T(0));
            if (same) res_b = finish_layer(Sb, A[1], 0, T(0));
            // layer 1: upper node of a's cell = upper (same cell) or lower (next cell) node of b's
            layer_contract<T, ND, 2>(tile_loader(lay + Cfg::STRIDE), W2, A);
            res_a = finish_layer(Sa, A[0], 1, res_a);
            res_b = same ? finish_layer(Sb, A[1], 1, res_b) : finish_layer(Sb, A[1], 0, T(0));
            if (!same) {
              // layer 2: upper node of b's cell
              WeightRefs<T, ND, 1> W1;
              W1.w[0] = Sb.w;
              T B[1][2];
              layer_contract<T, ND, 1>(tile_loader(lay + 2 * Cfg::STRIDE), W1, B);
              res_b = finish_layer(Sb, B[0], 1, res_b);
            }
            done_b = true;
          } else {
            WeightRefs<T, ND, 1> W1;
            W1.w[0] = Sa.w;
            T A[1][2];
            layer_contract<T, ND, 1>(tile_loader(lay), W1, A);
            res_a = finish_layer(Sa, A[0], 0, T(0));
            layer_contract<T, ND, 1>(tile_loader(lay + Cfg::STRIDE), W1, A);
            res_a = finish_layer(Sa, A[0], 1, res_a);
          }
        }
        __syncwarp();
        todo &= ~subm;
        if (todo) begin_pass();
      }
    } else if (act_a) {
      res_a = eval_direct(Sa, pza);
    }
    // second queries that did not ride (other column, or further than one cell away)
    if (!done_b) res_b = eval_direct(Sb, pzb);

    if (act_b && pair_store) {
      store_pair(Sa, Sb, qa, res_a, res_b);
    } else {
      if (act_a) store_q(Sa, qa, res_a);
      if (act_b) store_q(Sb, qa + 1, res_b);
    }
  }
#undef IPXT_WRAP
#undef IPXT_PERIODIC
#undef IPXT_PERM
}

// shared memory of the staged knots and reciprocal widths
template <typename T, int ND> static size_t tile_knot_bytes(const EvalArgs<T, ND>& a) {
  size_t knots = 0;
  for (int ax = 0; ax < ND; ++ax) knots += 2 * (size_t)a.ax[ax].nk * sizeof(T);
  return (knots + 15) & ~(size_t)15;
}

// warps of one CTA: as many as the register file holds, fewer if the shared memory left beside
// the knots does not hold that many per-warp tiles
// per-device launch constants, looked up once per device (immutable afterwards)
struct TileDevice {
  int dev, sms, smem_optin;
};
static TileDevice tile_device() {
  constexpr int kMaxDev = 64;
  static std::atomic<int> ready[kMaxDev];
  static int sms[kMaxDev], optin[kMaxDev];
  TileDevice d{0, 148, 48 * 1024};
  cudaGetDevice(&d.dev);
  if (d.dev < 0 || d.dev >= kMaxDev) return d;
  if (ready[d.dev].load(std::memory_order_acquire) == 0) {
    int s = 148, o = 48 * 1024;
    if (cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, d.dev) != cudaSuccess) s = 148;
    if (cudaDeviceGetAttribute(&o, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.dev) != cudaSuccess || o <= 0)
      o = 48 * 1024;
    sms[d.dev] = s;
    optin[d.dev] = o;
    ready[d.dev].store(1, std::memory_order_release);
  }
  d.sms = sms[d.dev];
  d.smem_optin = optin[d.dev];
  return d;
}
// experiment switches, read once per process (profiles/sweep_tile_modes.sh)
static const char* tile_env(int which) {
  static const char* const mode = getenv("IPX_TILE_MODE");
  static const char* const runtime_flags = getenv("IPX_TILE_RUNTIME_FLAGS");
  return which == 0 ? mode : runtime_flags;
}

template <typename T, int ND, int RC, int QPL> static int tile_warps(const EvalArgs<T, ND>& a) {
  const int lim = tile_device().smem_optin;
  const long long room = (long long)lim - 1024 - (long long)tile_knot_bytes<T, ND>(a);  // 1 KB: static
  const long long fit = room / TileCfg<T, ND, RC, QPL>::WARP_TILE_BYTES;
  const int maxw = tile_max_threads<T>() / 32;
  return fit >= maxw ? maxw : (fit < 0 ? 0 : (int)fit);
}

template <typename T, int ND, int RC, bool TWO, int WM, int PM = WM>
static bool launch_tile_wm(const EvalArgs<T, ND>& a, cudaStream_t s) {
  constexpr int QPL = TWO ? 2 : 1;
  const int warps = tile_warps<T, ND, RC, QPL>(a);
  if (warps < 4) return false;  // the generic kernel serves such knot vectors
  const size_t smem = tile_knot_bytes<T, ND>(a) + (size_t)warps * TileCfg<T, ND, RC, QPL>::WARP_TILE_BYTES;
  bool ord0 = a.dp == nullptr;
  for (int ax = 0; ax < ND && ord0; ++ax) ord0 = a.hp.axis[ax].derivative <= 0;
  void (*kern)(const EvalArgs<T, ND>);
  if constexpr (WM >= 0) {
    kern = ord0 ? eval_tile_kernel<T, ND, false, true, RC, TWO, WM, PM>
                : eval_tile_kernel<T, ND, false, false, RC, TWO, WM, PM>;
  } else {
    kern = a.dp ? eval_tile_kernel<T, ND, true, false, RC, TWO, -1>
                : (ord0 ? eval_tile_kernel<T, ND, false, true, RC, TWO, -1>
                        : eval_tile_kernel<T, ND, false, false, RC, TWO, -1>);
  }
  // the opt-in shared-memory size of a kernel is a per-device attribute: raise it when a launch
  // needs more than any earlier one on that device (never lowered), not on every launch
  const TileDevice td = tile_device();
  {
    constexpr int kMaxDev = 64;
    static std::atomic<int> granted[3][kMaxDev];  // per kernel variant of this instantiation
    const int variant = WM >= 0 ? (ord0 ? 1 : 0) : (a.dp ? 2 : (ord0 ? 1 : 0));
    const int slot = td.dev >= 0 && td.dev < kMaxDev ? td.dev : 0;
    if ((int)smem > granted[variant][slot].load(std::memory_order_relaxed) || td.dev != slot) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return false;
      granted[variant][slot].store((int)smem, std::memory_order_relaxed);
    }
  }
  const int sms = td.sms;
  const int threads = warps * 32;
  int64_t need = (a.nq + QPL * threads - 1) / (QPL * threads);
  int64_t grid = sms;  // persistent: one CTA per SM
  if (need < grid) grid = need;
  kern<<<(int)grid, threads, smem, s>>>(a);
  return true;
}

// The periodic axes as a compile-time mask when the call allows it (host-resident operands,
// no knot permutation, un-padded tables) and the mask is one of the common ones: none, all,
// or all but the last axis.  Anything else runs the kernel that reads the flags per query.
template <typename T, int ND, int RC, bool TWO = true>
static bool launch_tile_rc(const EvalArgs<T, ND>& a, cudaStream_t s) {
  int mask = 0, pmask = 0;
  bool ct = a.dp == nullptr && tile_env(1) == nullptr;
  for (int ax = 0; ax < ND; ++ax) {
    const AxisDesc<T>& A = a.ax[ax];
    ct = ct && A.perm == nullptr;
    mask |= (A.wrap != 0) << ax;
    pmask |= (A.periodic != 0) << ax;
  }
  constexpr int ALL = (1 << ND) - 1;
  constexpr int HEAD = (1 << (ND - 1)) - 1;  // every axis but the last
  if (ct && pmask == mask) {  // un-padded tables (Interpolator1D/2D/3D)
    if (mask == 0) return launch_tile_wm<T, ND, RC, TWO, 0>(a, s);
    if (mask == ALL) return launch_tile_wm<T, ND, RC, TWO, ALL>(a, s);
    if constexpr (ND > 1) {
      if (mask == HEAD) return launch_tile_wm<T, ND, RC, TWO, HEAD>(a, s);
    }
  } else if (ct && pmask == 0) {  // padded tables (functional form on periodic axes)
    if (mask == ALL) return launch_tile_wm<T, ND, RC, TWO, ALL, 0>(a, s);
    if constexpr (ND > 1) {
      if (mask == HEAD) return launch_tile_wm<T, ND, RC, TWO, HEAD, 0>(a, s);
    }
  }
  return launch_tile_wm<T, ND, RC, TWO, -1>(a, s);
}

// true if (T, ND) has a tile kernel: records of at least 16 bytes
template <typename T, int ND> constexpr bool tile_supported() {
  return (1 << ND) * (int)sizeof(T) >= 16;
}

// Launches the tile kernel if it applies (record >= 16 B, knot vectors fit in shared memory,
// query arrays 16-byte aligned); returns false if the caller should use the generic kernel.
template <typename T, int ND>
static bool launch_tile(const EvalArgs<T, ND>& a, cudaStream_t s) {
  if constexpr (!tile_supported<T, ND>()) {
    return false;
  } else {
    int total_knots = 0;
    for (int ax = 0; ax < ND; ++ax) {
      total_knots += a.ax[ax].nk;
      if ((reinterpret_cast<uintptr_t>(a.ax[ax].q) & 15) != 0) return false;
    }
    if (total_knots > kSmemKnotCap) return false;
    if ((reinterpret_cast<uintptr_t>(a.tab[0]) & 31) != 0) return false;  // 256-bit direct gathers
    // Queries per cell decide the shape of a warp's round (measured on 256^3 / 512^3 f64,
    // profiles/sweep_tile_modes.sh): from about three per cell up, two consecutive queries
    // per lane share node layers; below that a lane's pair is rarely in adjacent cells and a
    // warp's 64 queries span several runs, so one query per lane keeps every lane busy.  The
    // long run (32 nodes) pays off below about five queries per cell (a round spans more than
    // 16 nodes); denser streams keep the short one, which leaves more of the SM's memory to L1
    // for the direct gathers of incoherent (random-order) warps.
    double cells = 1.0;
    for (int ax = 0; ax < ND; ++ax) cells *= (double)(a.ax[ax].n > 1 ? a.ax[ax].n - 1 : 1);
    const bool two = (double)a.nq >= 3.0 * cells;
    const bool long_run =
        (double)a.nq < 5.0 * cells &&
        (two ? tile_warps<T, ND, 32, 2>(a) >= tile_warps<T, ND, 16, 2>(a)
             : tile_warps<T, ND, 32, 1>(a) >= tile_warps<T, ND, 16, 1>(a));
    const char* e = tile_env(0);  // experiment switch (profiles/sweep_tile_modes.sh)
    if (e != nullptr) {
      if (e[0] == 'a') return launch_tile_rc<T, ND, 16, true>(a, s);
      if (e[0] == 'b') return launch_tile_rc<T, ND, 16, false>(a, s);
      if (e[0] == 'c') return launch_tile_rc<T, ND, 32, true>(a, s);
      if (e[0] == 'd') return launch_tile_rc<T, ND, 32, false>(a, s);
    }
    if (long_run) return two ? launch_tile_rc<T, ND, 32, true>(a, s) : launch_tile_rc<T, ND, 32, false>(a, s);
    return two ? launch_tile_rc<T, ND, 16, true>(a, s) : launch_tile_rc<T, ND, 16, false>(a, s);
  }
}

}  // namespace ipx
