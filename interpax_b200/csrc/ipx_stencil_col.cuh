// ipx_stencil_col.cuh -- "column" shape of the evaluation straight from f (3-D, dense cell-sorted
// streams): whole table rows of the cell columns a CTA is working on are staged in shared memory
// by the TMA unit, and queries are re-paired by cell so that a lane reads the 4^3 neighbourhood
// ONCE for two queries.  Included by ipx_stencil.cu (uses its helpers).
//
// Why.  A tricubic query needs 64 operands (512 B in f64).  Delivering operands to registers
// costs one LSU wavefront per 128 B whatever the source, so one query per lane is bound at
// 4 clk per query per SM (0.47 of the HBM roofline on C4) before any arithmetic; the instruction
// stream of the group-by-column kernels (tile kernel on records, dense stencil kernel) is 13.8 /
// 15.3 warp instructions per query, of which only 5.5 are FP64.  This shape removes the per-warp
// grouping, copy loops and address arithmetic:
//
//   * The CTA owns a contiguous segment of the (cell-sorted) stream and walks it in WINDOWS.  A
//     producer warp locates 32 probe queries of the next stretch, picks the longest prefix whose
//     probes stay inside K consecutive cell columns (same first-axis cell, consecutive second-axis
//     cells), and stages the 4 x (K+3) table rows those columns read -- each row one
//     cp.async.bulk of the full last-axis extent -- into one of two row buffers, completing on an
//     mbarrier (full[b]); consumer warps hand the buffer back through free[b].
//   * The consumer warps take the window's rounds of 64 consecutive queries in turn, two
//     CONSECUTIVE queries per lane.  In a dense cell-sorted stream the second query of a lane sits
//     in the first one's cell or in the next cell up the last axis; the lane then reads FIVE
//     values per row -- 8-byte shared-memory loads at immediate offsets from one base register --
//     and applies them to both queries, the second with its four taps shifted by the cell
//     distance (a zero fifth tap).  Operand delivery per query drops from 64 to 40 loads and
//     there is no grouping, no copy loop and no address arithmetic per row.  A second query that
//     does not ride (other column, further away) takes a second pass of the same code.
//
// Queries the hot path does not cover -- a failed interval guess, queries more than a period out,
// cells outside the staged columns (an unsorted stream), NaN results (a non-finite table entry
// under the zero tap, or a NaN query) -- go through the one-query path (st_redo), so results never
// depend on the stream order or the neighbours: the arithmetic of a query is st_eval_one's,
// operation for operation (bitwise equal, tested).  Irregular cells (st_stage) and derivative
// orders above 1 use the general taps (st_taps_general_inl) in line.
#pragma once
#include <type_traits>

namespace ipx {

#ifndef IPX_COL_WARPS
#define IPX_COL_WARPS 15
#endif
constexpr int kColWarps = IPX_COL_WARPS;   // consumer warps; one more warp produces
constexpr int kColStep = 320;   // spacing of the producer's probes: a window is at most 31 steps, whole rounds
constexpr int kColMaxSegs = 8;  // stream segments per CTA (interleaved over the CTAs)

template <typename T> struct ColCfg {
  static constexpr int XCH_BYTES = 3 * 64 * (int)sizeof(T);  // per warp: the next round's coordinates
};

// one cell of one axis as phase 1 reads it: lower knot and signed reciprocal width (st_stage)
template <typename T> struct alignas(2 * sizeof(T)) ColCell {
  T l, r;
};

struct ColHdr {
  int s, e;      // the window's queries [s, e), relative to the segment; s == e: no more windows
  int ix0, iy0;  // first cell column
  int ncol;      // columns staged: second-axis cells iy0 .. iy0 + ncol - 1
  int seg;       // segment index
  int pad[2];
};

__device__ __forceinline__ bool col_mbar_test(unsigned mbar_s, unsigned phase) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(mbar_s), "r"(phase)
      : "memory");
  return ok != 0;
}
// blocking wait: try_wait with a suspend-time hint parks the warp in hardware until the phase
// completes (or the hint, about 10 ms, runs out) instead of polling
__device__ __forceinline__ void col_mbar_wait(unsigned mbar_s, unsigned phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "IPX_COL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra IPX_COL_DONE;\n"
      "bra IPX_COL_WAIT;\n"
      "IPX_COL_DONE:\n"
      "}\n" ::"r"(mbar_s), "r"(phase), "r"(0x989680u) : "memory");
}
// "buffer b is free again": a named barrier (ids 1, 2) per row buffer.  Consumer warps arrive when
// they are done with a window, the producer warp syncs before it refills the buffer -- blocked in
// hardware, where polling an mbarrier cost 6 % of the kernel's instructions on the scheduler the
// producer shares with three consumers (try_wait's suspend hint parks a warp for ~20 ns at a time).
__device__ __forceinline__ void col_free_arrive(int b) {
  asm volatile("bar.arrive %0, %1;" ::"r"(1 + b), "r"((kColWarps + 1) * 32) : "memory");
}
__device__ __forceinline__ void col_free_wait(int b) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + b), "r"((kColWarps + 1) * 32) : "memory");
}
__device__ __forceinline__ void col_mbar_arrive(unsigned mbar_s) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(mbar_s) : "memory");
}
template <int BYTES> __device__ __forceinline__ void col_cp_async(unsigned smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_dst), "l"(gmem_src), "n"(BYTES) : "memory");
}

// taps of a regular cell from the local coordinate (st_taps_regular with t already formed)
template <typename T, bool ORD0> __device__ __forceinline__ void col_taps(T t, T r, int order, T (&tp)[4]) {
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

__device__ __forceinline__ void col_lds_pair(const double* p, double& x, double& y) {
  const double2 v = *reinterpret_cast<const double2*>(p);
  x = v.x; y = v.y;
}
__device__ __forceinline__ void col_lds_pair(const float* p, float& x, float& y) {
  const float2 v = *reinterpret_cast<const float2*>(p);
  x = v.x; y = v.y;
}
// 0 <= t < 1 - 2^-21 (float: 1 - 2^-8), tested on the integer pipe: NaN, negative values and -0.0
// fail it.  Passing it proves l <= v < h for a cell of non-zero width (t = (v - l) * |1 / (h - l)|).
__device__ __forceinline__ bool col_inside(double t) { return (unsigned)__double2hiint(t) < 0x3fefffffu; }
__device__ __forceinline__ bool col_inside(float t) { return (unsigned)__float_as_int(t) < 0x3f7f0000u; }

// shared memory after the knot tables: [cells of the 3 axes][exchange buffers][row buffer 0][row buffer 1]
// WM: set of axes whose queries are wrapped, as a compile-time mask (-1: read the flags).
// registers per thread: what one CTA of kColWarps + 1 warps per SM leaves (multiple of 8)
constexpr int kColRegs = (65536 / ((kColWarps + 1) * 32)) / 8 * 8 > 255 ? 255 : (65536 / ((kColWarps + 1) * 32)) / 8 * 8;
template <typename T, bool DEVP, bool ORD0, int WM>
__global__ void __maxnreg__(kColRegs)
stencil_column_kernel(const __grid_constant__ StArgs<T, 3> a, const int ry, const int rs, const int nseg, const int seg) {
  constexpr int ND = 3;
  using Cfg = ColCfg<T>;
#define IPXC_WRAP(AXI) (WM >= 0 ? (((WM >> (AXI)) & 1) != 0) : (a.ax[AXI].wrap != 0))
  extern __shared__ __align__(32) char smem[];
  __shared__ StAxC<T> AXC[ND];
  __shared__ ipx_params params_s;
  __shared__ __align__(16) ColHdr HDR[2];
  __shared__ __align__(8) unsigned long long MBAR[2];  // full[0], full[1] (rows of a window have landed)
  __shared__ int NEXT_ROUND[2];  // per row buffer: the next round of its window nobody has taken yet
  if (a.flag != nullptr && ((a.accept >> (*a.flag & 31)) & 1u) == 0) return;
  if (DEVP && threadIdx.x == 32) params_s = *a.dp;
  const ipx_params& P = DEVP ? params_s : a.hp;
  const unsigned mbar_s = (unsigned)__cvta_generic_to_shared(MBAR);
  if (threadIdx.x == 0) {
    st_mbar_init(mbar_s, 1);
    st_mbar_init(mbar_s + 8, 1);
  }
  StShared<T, ND> S;
  st_stage<T, ND>(a, smem, AXC, S);  // (ends with a barrier: the mbarriers are initialised too)
  const int tables = st_layout<T, ND>(a).total;
  // per axis, cell i as one 16-byte (float: 8-byte) entry {ks[i-1], rr[i]}
  {
    int off = tables;
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      ColCell<T>* cl = reinterpret_cast<ColCell<T>*>(smem + off);
      const int nk = a.ax[ax].nk;
      for (int i = threadIdx.x; i <= nk; i += (int)blockDim.x) {  // (entry nk: the last knot, for cell nk - 1)
        ColCell<T> c;
        c.l = i > 0 ? S.ks[ax][i - 1] : T(0);
        c.r = i < nk ? S.rr[ax][i] : T(0);
        cl[i] = c;
      }
      off += ((nk + 1) * (int)sizeof(ColCell<T>) + 15) & ~15;
    }
  }
  __syncthreads();
  int cells_bytes = 0;
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) cells_bytes += ((a.ax[ax].nk + 1) * (int)sizeof(ColCell<T>) + 15) & ~15;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  char* const xch0 = smem + tables + cells_bytes;
  char* const rows0 = xch0 + kColWarps * Cfg::XCH_BYTES;
  const int buf_bytes = 4 * ry * rs;
  const int row_bytes = (int)a.pitch * (int)sizeof(T);
  const int nky = a.ax[1].nk;

  if (warp == kColWarps) {
    // ------------------------------- producer ---------------------------------------
    const int K = ry - 3;
    int k = 0;
    for (int sg = 0; sg < nseg; ++sg) {
      // the segments are dealt to the CTAs in turn (neighbouring CTAs share table rows in L2)
      const int64_t seg0 = ((int64_t)sg * gridDim.x + blockIdx.x) * seg;
      const int len = (int)max((int64_t)0, min((int64_t)seg, a.nq - seg0));
      int s = 0;
      while (s < len) {
        const int b = k & 1;
        // probes: every kColStep-th query of the stretch, located on the two column axes
        const int pos = min(s + lane * kColStep, len - 1);
        int pi[2];
#pragma unroll
        for (int ax = 0; ax < 2; ++ax) {
          T v = a.ax[ax].q[seg0 + pos];
          if (IPXC_WRAP(ax)) v = py_mod_slow(v, T(fabs(P.axis[ax].period)));
          pi[ax] = bin_index_slow(S.ks[ax], a.ax[ax].nk, v, AXC[ax].x_first, AXC[ax].inv_h);
        }
        const int ix0 = __shfl_sync(0xffffffffu, pi[0], 0);
        const int iy0 = __shfl_sync(0xffffffffu, pi[1], 0);
        const bool inside = pi[0] == ix0 && pi[1] >= iy0 && pi[1] < iy0 + K;
        const unsigned m = __ballot_sync(0xffffffffu, inside);
        const int lead = m == 0xffffffffu ? 32 : __ffs(~m) - 1;  // >= 1: lane 0 defines the column
        const int iy_last = __shfl_sync(0xffffffffu, pi[1], lead - 1);
        const int pos_last = __shfl_sync(0xffffffffu, pos, lead - 1);
        int e, ncol;
        if (pos_last >= len - 1) {
          e = len;  // the rest of the segment
          ncol = iy_last - iy0 + 1;
        } else if (lead == 1) {
          // not even the second probe is inside (sparse or unsorted stretch): take one step anyway;
          // what falls outside the staged columns goes through the one-query path
          e = min(s + kColStep, len);
          ncol = min(K, nky - iy0);
        } else {
          e = pos_last;  // (whole rounds: the last probe inside starts the next window)
          ncol = iy_last - iy0 + 1;
        }
        // (the probes did not need the buffer; now wait until the consumers have handed it back,
        // parked in hardware: this warp shares its scheduler with three consumer warps)
        if (k >= 2) col_free_wait(b);
        const int nry = ncol + 3;
        const int nrows = 4 * nry;
        if (lane == 0) {
          HDR[b].s = s; HDR[b].e = e; HDR[b].ix0 = ix0; HDR[b].iy0 = iy0; HDR[b].ncol = ncol; HDR[b].seg = sg;
          NEXT_ROUND[b] = 0;
          st_mbar_expect_tx(mbar_s + 8 * b, (unsigned)(nrows * row_bytes));
        }
        __syncwarp();
        const unsigned dst0 = (unsigned)__cvta_generic_to_shared(rows0 + b * buf_bytes);
        const T* src0 = a.tab + (int64_t)(ix0 - 1) * a.ax[0].stride + (int64_t)(iy0 - 1) * a.ax[1].stride;
        for (int r = lane; r < nrows; r += 32) {
          const int pa = r / nry, yy = r - pa * nry;
          st_bulk_g2s(dst0 + (unsigned)((pa * ry + yy) * rs),
                      src0 + (int64_t)pa * a.ax[0].stride + (int64_t)yy * a.ax[1].stride, (unsigned)row_bytes,
                      mbar_s + 8 * b);
        }
        s = e;
        ++k;
      }
    }
    // no more windows
    const int b = k & 1;
    if (k >= 2) col_free_wait(b);
    if (lane == 0) {
      HDR[b].s = 0; HDR[b].e = 0;
      col_mbar_arrive(mbar_s + 8 * b);
    }
    return;
  }

  // --------------------------------- consumers ----------------------------------------
  char* const xch = xch0 + warp * Cfg::XCH_BYTES;
  const T* const cb = reinterpret_cast<const T*>(xch) + 2 * lane;  // the lane's two slots of axis 0
  const unsigned cb_s = (unsigned)__cvta_generic_to_shared(xch) + (unsigned)(2 * lane * (int)sizeof(T));
  bool can_fill[ND];
  unsigned high_order = 0;  // axes whose derivative order the regular-cell taps do not cover
  bool vec_ok = true;       // 16-byte (float: 8-byte) copies of a lane's two coordinates
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    can_fill[ax] = !IPXC_WRAP(ax) && (!P.axis[ax].keep_lo || !P.axis[ax].keep_hi);
    if (!ORD0 && P.axis[ax].derivative > 1) high_order |= 1u << ax;
    vec_ok = vec_ok && (reinterpret_cast<uintptr_t>(a.ax[ax].q) & (2 * sizeof(T) - 1)) == 0;
  }
  const int plane_bytes = ry * rs;
  const bool out_vec_ok = (reinterpret_cast<uintptr_t>(a.out) & (2 * sizeof(T) - 1)) == 0;
  // general taps of cell i of one axis; the table pointers are formed here, behind an opaque zero,
  // so that the compiler does not carry nine of them through the hot loop for a rare branch
  auto general_taps = [&](int ax, int i, T t, T (&tp)[4]) {
    int zero;
    asm volatile("mov.u32 %0, 0;" : "=r"(zero));
    const StLayout<ND> lay = st_layout<T, ND>(a);
    const char* sm = smem + zero;
    const T* ks = reinterpret_cast<const T*>(sm + lay.ks_off[ax]);
    const T* rr = reinterpret_cast<const T*>(sm + lay.rr_off[ax]);
    const Coef2<T>* ac = reinterpret_cast<const Coef2<T>*>(sm + lay.ac_off[ax]);
    const Taps4<T> g = st_taps_general_inl<T>(ks, ac, i, t, t_abs(rr[i]), ORD0 ? 0 : P.axis[ax].derivative);
    tp[0] = g.t0; tp[1] = g.t1; tp[2] = g.t2; tp[3] = g.t3;
  };

  // the coordinates of the lane's two queries of the round at `pos` travel global -> shared (its
  // own slots, entries 2 * lane and 2 * lane + 1 of each axis) while the current round is computed
  auto prefetch = [&](const int64_t seg0, int pos) {
    const int64_t qi = seg0 + pos + 2 * lane;
    if (vec_ok && qi + 1 < a.nq) {
#pragma unroll
      for (int ax = 0; ax < ND; ++ax)
        col_cp_async<2 * (int)sizeof(T)>(cb_s + (unsigned)(ax * 64 * (int)sizeof(T)), a.ax[ax].q + qi);
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        if (qi + k < a.nq) {
#pragma unroll
          for (int ax = 0; ax < ND; ++ax)
            col_cp_async<(int)sizeof(T)>(cb_s + (unsigned)((ax * 64 + k) * (int)sizeof(T)), a.ax[ax].q + qi + k);
        }
    }
  };

  // The warps TAKE the rounds of a window (an atomic counter in shared memory) rather than being
  // dealt them: the scheduler that hosts the producer warp has one consumer warp fewer than the
  // others, and with equal shares its warps ran ahead and waited at every window.
  auto take_round = [&](int buf) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&NEXT_ROUND[buf], 1);
    return __shfl_sync(0xffffffffu, r, 0);
  };
  int carried = -1;  // a round of the NEXT window taken (and its coordinates requested) already
  for (int k = 0;; ++k) {
    const int b = k & 1;
    col_mbar_wait(mbar_s + 8 * b, (unsigned)((k >> 1) & 1));
    const int4 hd = *reinterpret_cast<const int4*>(&HDR[b]);
    const int ws = hd.x, we = hd.y;
    if (ws >= we) break;
    const int nrounds = (we - ws + 63) >> 6;
    int rnd = carried;
    if (rnd < 0) {
      rnd = take_round(b);
      if (rnd < nrounds) prefetch(((int64_t)HDR[b].seg * gridDim.x + blockIdx.x) * seg, ws + 64 * rnd);
    }
    carried = -1;
    int nrnd = nrounds;  // the round after rnd (taken while rnd is worked on)

    // One round.  FAST: the round as the hot loop runs it, free of calls (out-of-line calls in
    // rare branches make the compiler spill around them in the loop); it gives up -- results it
    // has stored are stored again -- if a query of the round needs the one-query path, and the
    // round is then redone by the other instantiation (coordinates straight from global memory).
    auto round = [&](auto fast_tag) -> bool {
      constexpr bool FAST = decltype(fast_tag)::value;
      // (the window's constants are read from shared memory every round: registers are what the
      // contraction below is short of)
      // (likewise the cell tables' addresses: re-formed behind an opaque zero instead of carried --
      // and spilled -- through the loop)
      const ColCell<T>* CELL[ND];
      {
        int zero;
        asm volatile("mov.u32 %0, 0;" : "=r"(zero));
        int off = st_layout<T, ND>(a).total + zero;
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) {
          CELL[ax] = reinterpret_cast<const ColCell<T>*>(smem + off);
          off += ((a.ax[ax].nk + 1) * (int)sizeof(ColCell<T>) + 15) & ~15;
        }
      }
      const int4 hw = *reinterpret_cast<const int4*>(&HDR[b]);
      const int ws = hw.x, we = hw.y, ix0 = hw.z, iy0 = hw.w, ncol = HDR[b].ncol;
      const int64_t seg0 = ((int64_t)HDR[b].seg * gridDim.x + blockIdx.x) * seg;
      const char* const rows = rows0 + b * buf_bytes;
      T* const outs = a.out + seg0;
      const ColCell<T> cell_x = CELL[0][ix0];
      const T cell_hx = CELL[0][ix0 + 1].l;
      const unsigned flags0 = (high_order << 6) | (((unsigned)sign_word(cell_x.r) >> 31) << 6);
      const int pos = ws + 64 * rnd;
      const int qa = pos + 2 * lane;  // the lane's queries: qa and qa + 1 (of the segment)
      T c[ND][2];
      if constexpr (FAST) {
        st_cp_async_wait();
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) col_lds_pair(cb + ax * 64, c[ax][0], c[ax][1]);
        // (the lane's own slots, read above: the next round's coordinates can follow at once)
        nrnd = take_round(b);
        if (nrnd < nrounds) {
          prefetch(seg0, ws + 64 * nrnd);
        } else {
          // nothing left in this window: start on the next one if its header is already there
          const int nb = b ^ 1;
          if (col_mbar_test(mbar_s + 8 * nb, (unsigned)(((k + 1) >> 1) & 1))) {
            const int ns = HDR[nb].s, ne = HDR[nb].e;
            if (ns < ne) {
              carried = take_round(nb);
              if (ns + 64 * carried < ne)
                prefetch(((int64_t)HDR[nb].seg * gridDim.x + blockIdx.x) * seg, ns + 64 * carried);
            }
          }
        }
      } else {
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int ax = 0; ax < ND; ++ax) c[ax][kk] = qa + kk < we ? a.ax[ax].q[seg0 + qa + kk] : T(0);
      }
      // ---- phase 1: both queries located ----
      unsigned flg[2];  // bits 0-5 fill, bits 6-8 axes that need the general taps
      bool slow[2];
      int iy[2], iz[2];
      T tt[2][ND];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        bool sl = false;
        unsigned fl = flags0;
        int ii[ND];
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) {
          const int nk = a.ax[ax].nk;
          T v = c[ax][kk];
          if (IPXC_WRAP(ax)) {
            bool wrap_bad;
            v = wrap_one_period(v, T(fabs(P.axis[ax].period)), wrap_bad);
            sl = sl || wrap_bad;
          }
          // first axis: every cell of the window is ix0 (anything else fails the test below);
          // the others: guess from the uniform spacing
          const int i = ax == 0 ? ix0 : min(max(t_floor_to_int((v - AXC[ax].x_first) * AXC[ax].inv_h), 0), nk - 2) + 1;
          const ColCell<T> cl = ax == 0 ? cell_x : CELL[ax][i];
          const T t = (v - cl.l) * t_abs(cl.r);
          const unsigned irregular = (unsigned)sign_word(cl.r) >> 31;
          if (ax > 0) fl |= irregular << (6 + ax);
          if (!col_inside(t) || irregular) {
            // near a knot, outside the knots, NaN, or a cell that may have zero width: the exact test
            const T h = ax == 0 ? cell_hx : CELL[ax][i + 1].l;
            if (!((v >= cl.l || i == 1) && (v < h || i == nk - 1))) sl = true;
            if (can_fill[ax]) fl |= st_fill_bits(P.axis[ax], AXC[ax], i, nk, v) << (2 * ax);
          }
          tt[kk][ax] = t;
          ii[ax] = i;
        }
        sl = sl || (unsigned)(ii[1] - iy0) >= (unsigned)ncol;
        slow[kk] = sl && qa + kk < we;
        flg[kk] = fl;
        iy[kk] = ii[1];
        iz[kk] = ii[2];
      }
      if constexpr (FAST) {
        if (__any_sync(0xffffffffu, slow[0] || slow[1])) return false;
      }
      const bool val0 = qa < we, val1 = qa + 1 < we;
      // the second query rides with the first if it sits in its cell or the next one up the last axis
      const int dz = iz[1] - iz[0];
      const bool ride = val1 && !slow[0] && !slow[1] && iy[1] == iy[0] && (dz == 0 || dz == 1);
      // second queries that do not ride (other column, more than a cell away: rare in a dense
      // cell-sorted stream) take a second pass of the same code -- in the cold instantiation only
      bool pend = val1 && !slow[1] && !ride;
      if constexpr (FAST) {
        if (__any_sync(0xffffffffu, pend)) return false;
      }
      bool nan_seen = false;
      // the pass's first query (no run-time indexing: the arrays stay in registers)
      T at[ND];
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) at[ax] = tt[0][ax];
      unsigned aflg = flg[0];
      int cy = iy[0], cz = iz[0], ka = 0;
      bool act_a = val0 && !slow[0], act_b = ride;
#pragma unroll 1
      for (int pass = 0; pass < (FAST ? 1 : 2); ++pass) {
        bool nan_a = false, nan_b = false;
        if (act_a) {
          T TA[ND][4], TBX[4], TBY[4], TBZ[5];
          const unsigned gen_a = (aflg >> 6) & 7u, gen_b = (flg[1] >> 6) & 7u;
#pragma unroll
          for (int ax = 0; ax < ND; ++ax) {
            const int i = ax == 0 ? ix0 : (ax == 1 ? cy : cz);
            if ((gen_a >> ax) & 1u) {
              general_taps(ax, i, at[ax], TA[ax]);
            } else {
              col_taps<T, ORD0>(at[ax], ORD0 ? T(0) : t_abs(CELL[ax][i].r), P.axis[ax].derivative, TA[ax]);
            }
          }
          {
            // the rider's taps: its own cell along the last axis (iz[1]), the pair's on the others
            T tb[ND][4];
#pragma unroll
            for (int ax = 0; ax < ND; ++ax) {
              const int i = ax == 0 ? ix0 : (ax == 1 ? cy : iz[1]);
              if (act_b && ((gen_b >> ax) & 1u)) {
                general_taps(ax, i, tt[1][ax], tb[ax]);
              } else {
                col_taps<T, ORD0>(tt[1][ax], ORD0 ? T(0) : t_abs(CELL[ax][i].r), P.axis[ax].derivative, tb[ax]);
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { TBX[j] = tb[0][j]; TBY[j] = tb[1][j]; }
            const bool same = dz == 0;
            TBZ[0] = same ? tb[2][0] : T(0);
            TBZ[1] = same ? tb[2][1] : tb[2][0];
            TBZ[2] = same ? tb[2][2] : tb[2][1];
            TBZ[3] = same ? tb[2][3] : tb[2][2];
            TBZ[4] = same ? T(0) : tb[2][3];
          }
          const char* base = rows + (cy - iy0) * rs + (cz - 1) * (int)sizeof(T);
          T acc_a = T(0), acc_b = T(0);
#pragma unroll
          for (int aa = 0; aa < 4; ++aa) {
            T sya = T(0), syb = T(0);
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
              const T* p = reinterpret_cast<const T*>(base + bb * rs);
              const T f0 = p[0], f1 = p[1], f2 = p[2], f3 = p[3], f4 = p[4];
              T ra = TA[2][0] * f0, rb = TBZ[0] * f0;
              ra = fma(TA[2][1], f1, ra); rb = fma(TBZ[1], f1, rb);
              ra = fma(TA[2][2], f2, ra); rb = fma(TBZ[2], f2, rb);
              ra = fma(TA[2][3], f3, ra); rb = fma(TBZ[3], f3, rb);
              rb = fma(TBZ[4], f4, rb);
              sya = fma(TA[1][bb], ra, sya);
              syb = fma(TBY[bb], rb, syb);
            }
            acc_a = fma(TA[0][aa], sya, acc_a);
            acc_b = fma(TBX[aa], syb, acc_b);
            base += plane_bytes;
          }
          // a NaN may be a non-finite table entry under the zero tap: the one-query path decides
          nan_a = acc_a != acc_a;
          nan_b = act_b && acc_b != acc_b;
          const T res_a = st_apply_fill<T, ND>(P, aflg & 63u, acc_a);
          if (act_b) {
            // (qa is even and the segments start at multiples of 64: the pair is aligned if `out` is)
            const T res_b = st_apply_fill<T, ND>(P, flg[1] & 63u, acc_b);
            if (out_vec_ok) {
              if constexpr (sizeof(T) == 8) {
                __stcs(reinterpret_cast<double2*>(outs + qa), make_double2(res_a, res_b));
              } else {
                __stcs(reinterpret_cast<float2*>(outs + qa), make_float2(res_a, res_b));
              }
            } else {
              __stcs(outs + qa, res_a);
              __stcs(outs + qa + 1, res_b);
            }
          } else {
            __stcs(outs + qa + ka, res_a);
          }
        }
        nan_seen = nan_seen || nan_a || nan_b;
        if constexpr (!FAST) {
          if (nan_a) { if (ka == 0) slow[0] = true; else slow[1] = true; }
          if (nan_b) slow[1] = true;
          if (!__any_sync(0xffffffffu, pend)) break;
          // second pass: the second queries that did not ride, alone
#pragma unroll
          for (int ax = 0; ax < ND; ++ax) at[ax] = tt[1][ax];
          aflg = flg[1]; cy = iy[1]; cz = iz[1]; ka = 1;
          act_a = pend; act_b = false; pend = false;
        }
      }
      if constexpr (FAST) {
        if (__any_sync(0xffffffffu, nan_seen)) return false;
      } else {
        if (slow[0] || slow[1]) {
          if (slow[0]) st_redo<T, ND, ORD0>(&a, &P, smem, AXC, seg0 + qa, 1u);
          if (slow[1]) st_redo<T, ND, ORD0>(&a, &P, smem, AXC, seg0 + qa + 1, 1u);
        }
      }
      return true;
    };
    while (rnd < nrounds) {
      if (!round(std::true_type{})) round(std::false_type{});
      rnd = nrnd;
    }
    __syncwarp();
    col_free_arrive(b);
  }
#undef IPXC_WRAP
}

// shared-memory plan of a launch; false when the shape does not apply (long knot vectors or rows,
// interval indices requested)
struct ColPlan {
  long long ry, rs;
  size_t smem;
};
template <typename T> static bool st_column_plan(const StArgs<T, 3>& a, const StDevice& d, ColPlan& pl) {
  const StLayout<3> lay = st_layout<T, 3>(a);
  long long cells = 0;
  for (int ax = 0; ax < 3; ++ax) {
    if (a.ax[ax].nk > 1023) return false;  // cell indices travel in 10 bits each
    cells += ((long long)(a.ax[ax].nk + 1) * (long long)sizeof(ColCell<T>) + 15) & ~15ll;
  }
  if (a.idx_out != nullptr) return false;
  pl.rs = (long long)a.pitch * (long long)sizeof(T);  // row stride in shared memory = row bytes
  const long long fixed = (long long)lay.total + cells + (long long)kColWarps * ColCfg<T>::XCH_BYTES;
  const long long room = (long long)d.smem_optin - 2048 - fixed;
  pl.ry = room / (2 * 4 * pl.rs);
  if (pl.ry > 32) pl.ry = 32;
  if (pl.ry < 5) return false;  // fewer than two columns per window
  pl.smem = (size_t)(fixed + 2 * 4 * pl.ry * pl.rs);
  return true;
}

// launch; IPX_EUNSUPPORTED when the shape does not apply (the caller uses another one)
template <typename T>
static int st_launch_column(const StArgs<T, 3>& a, const StDevice& d, cudaStream_t s) {
  ColPlan pl;
  if (!st_column_plan<T>(a, d, pl)) return IPX_EUNSUPPORTED;
  const long long ry = pl.ry, rs = pl.rs;
  const size_t smem = pl.smem;
  bool ord0 = a.dp == nullptr;
  for (int ax = 0; ax < 3 && ord0; ++ax) ord0 = a.hp.axis[ax].derivative <= 0;
  int mask = 0;
  for (int ax = 0; ax < 3; ++ax) mask |= (a.ax[ax].wrap != 0) << ax;
  void (*kern)(const StArgs<T, 3>, int, int, int, int);
  if (a.dp) kern = stencil_column_kernel<T, true, false, -1>;
  else if (mask == 0) kern = ord0 ? stencil_column_kernel<T, false, true, 0> : stencil_column_kernel<T, false, false, 0>;
  else if (mask == 3) kern = ord0 ? stencil_column_kernel<T, false, true, 3> : stencil_column_kernel<T, false, false, 3>;
  else kern = ord0 ? stencil_column_kernel<T, false, true, -1> : stencil_column_kernel<T, false, false, -1>;
  if (!st_grant_smem(reinterpret_cast<const void*>(kern), d.dev, smem)) return IPX_ELAUNCH;
  int64_t need = (a.nq + 2047) / 2048;
  int64_t grid = d.sms;  // persistent: one CTA per SM
  if (need < grid) grid = need;
  // the stream in grid * nseg segments, dealt to the CTAs in turn: a stretch of expensive cells
  // (the seam planes of a periodic axis take the general taps) is shared by several CTAs
  int64_t nseg = a.nq / (grid * 65536);
  nseg = nseg < 1 ? 1 : (nseg > kColMaxSegs ? kColMaxSegs : nseg);
  int64_t seg = (a.nq + grid * nseg - 1) / (grid * nseg);
  seg = (seg + 63) & ~(int64_t)63;
  if (seg >= (1ll << 30)) return IPX_EUNSUPPORTED;
  kern<<<(int)grid, (kColWarps + 1) * 32, smem, s>>>(a, (int)ry, (int)rs, (int)nseg, (int)seg);
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

}  // namespace ipx
