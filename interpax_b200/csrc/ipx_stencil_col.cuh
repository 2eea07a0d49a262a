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
//   * A consumer warp takes a contiguous share of the window, 64 queries per round (two per lane),
//     locates them (phase 1), and RE-PAIRS them: consecutive queries of one cell form a pair, a
//     pair takes a lane (a cell with an odd count leaves one half empty).  Local coordinates and
//     cell indices travel through a 2 KB per-warp exchange buffer; queries that do not fit the
//     32 lanes are simply read again next round (the cursor advances by what was taken).
//   * Phase 2: each lane forms the taps of its pair and walks the 16 rows of the cell: four
//     8-byte shared-memory loads at immediate offsets from one base register per row, applied to
//     both queries.  Operand delivery per query halves; there is no address arithmetic per row.
//
// Queries the hot path does not cover -- irregular cells (st_stage), derivative orders above 1,
// a failed interval guess, cells outside the staged columns (an unsorted stream) -- go through
// the one-query path (st_redo), so results never depend on the stream order or the neighbours:
// the arithmetic of a query is st_eval_one's, operation for operation (bitwise equal, tested).
#pragma once
#include <type_traits>

namespace ipx {

#ifndef IPX_COL_WARPS
#define IPX_COL_WARPS 15
#endif
constexpr int kColWarps = IPX_COL_WARPS;   // consumer warps; one more warp produces
constexpr int kColSub = 640;    // most queries per consumer warp and window
constexpr int kColStep = kColWarps * kColSub / 32;  // spacing of the producer's probes
constexpr int kColMaxSegs = 8;  // stream segments per CTA (interleaved over the CTAs)

template <typename T> struct ColCfg {
  static constexpr int TQ_BYTES = 3 * 64 * (int)sizeof(T);   // local coordinates of 64 entries
  static constexpr int XCH_BYTES = TQ_BYTES + 64 * 8;        // + 8 bytes of cell / flags each
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
template <typename T, bool DEVP, bool ORD0, int WM>
__global__ void __launch_bounds__((kColWarps + 1) * 32, 1)
stencil_column_kernel(const __grid_constant__ StArgs<T, 3> a, const int ry, const int rs, const int nseg, const int seg) {
  constexpr int ND = 3;
  using Cfg = ColCfg<T>;
#define IPXC_WRAP(AXI) (WM >= 0 ? (((WM >> (AXI)) & 1) != 0) : (a.ax[AXI].wrap != 0))
  extern __shared__ __align__(32) char smem[];
  __shared__ StAxC<T> AXC[ND];
  __shared__ ipx_params params_s;
  __shared__ __align__(16) ColHdr HDR[2];
  __shared__ __align__(8) unsigned long long MBAR[4];  // full[0], full[1], free[0], free[1]
  if (a.flag != nullptr && ((a.accept >> (*a.flag & 31)) & 1u) == 0) return;
  if (DEVP && threadIdx.x == 32) params_s = *a.dp;
  const ipx_params& P = DEVP ? params_s : a.hp;
  const unsigned mbar_s = (unsigned)__cvta_generic_to_shared(MBAR);
  if (threadIdx.x == 0) {
    st_mbar_init(mbar_s, 1);
    st_mbar_init(mbar_s + 8, 1);
    st_mbar_init(mbar_s + 16, kColWarps);
    st_mbar_init(mbar_s + 24, kColWarps);
  }
  StShared<T, ND> S;
  st_stage<T, ND>(a, smem, AXC, S);  // (ends with a barrier: the mbarriers are initialised too)
  const int tables = st_layout<T, ND>(a).total;
  // per axis, cell i as one 16-byte (float: 8-byte) entry {ks[i-1], rr[i]}
  const ColCell<T>* CELL[ND];
  {
    int off = tables;
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      ColCell<T>* cl = reinterpret_cast<ColCell<T>*>(smem + off);
      CELL[ax] = cl;
      const int nk = a.ax[ax].nk;
      for (int i = threadIdx.x; i < nk; i += (int)blockDim.x) {
        ColCell<T> c;
        c.l = i > 0 ? S.ks[ax][i - 1] : T(0);
        c.r = S.rr[ax][i];
        cl[i] = c;
      }
      off += (nk * (int)sizeof(ColCell<T>) + 15) & ~15;
    }
  }
  __syncthreads();
  int cells_bytes = 0;
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) cells_bytes += (a.ax[ax].nk * (int)sizeof(ColCell<T>) + 15) & ~15;

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
        if (lead == 1 && pos_last < len - 1) {
          // not even the second probe is inside (sparse or unsorted stretch): take one step anyway;
          // what falls outside the staged columns goes through the one-query path
          e = min(s + kColStep, len);
          ncol = min(K, nky - iy0);
        } else {
          e = pos_last + 1;
          ncol = iy_last - iy0 + 1;
        }
        // (the probes did not need the buffer; now wait until the consumers have handed it back,
        // sleeping between polls: this warp shares its scheduler with three consumer warps)
        if (k >= 2) {
          while (!col_mbar_test(mbar_s + 16 + 8 * b, (unsigned)(((k >> 1) + 1) & 1))) __nanosleep(200);
        }
        const int nry = ncol + 3;
        const int nrows = 4 * nry;
        if (lane == 0) {
          HDR[b].s = s; HDR[b].e = e; HDR[b].ix0 = ix0; HDR[b].iy0 = iy0; HDR[b].ncol = ncol; HDR[b].seg = sg;
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
    if (k >= 2) {
      while (!col_mbar_test(mbar_s + 16 + 8 * b, (unsigned)(((k >> 1) + 1) & 1))) __nanosleep(200);
    }
    if (lane == 0) {
      HDR[b].s = 0; HDR[b].e = 0;
      col_mbar_arrive(mbar_s + 8 * b);
    }
    return;
  }

  // --------------------------------- consumers ----------------------------------------
  char* const xch = xch0 + warp * Cfg::XCH_BYTES;
  T* const tq = reinterpret_cast<T*>(xch);                            // [3][64]
  uint2* const meta = reinterpret_cast<uint2*>(xch + Cfg::TQ_BYTES);  // [64]
  const unsigned xch_s = (unsigned)__cvta_generic_to_shared(xch) + (unsigned)(2 * lane * (int)sizeof(T));
  const unsigned le_mask = 0xffffffffu >> (31 - lane);
  bool can_fill[ND];
  unsigned high_order = 0;  // axes whose derivative order the regular-cell taps do not cover
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) {
    can_fill[ax] = !IPXC_WRAP(ax) && (!P.axis[ax].keep_lo || !P.axis[ax].keep_hi);
    if (!ORD0 && P.axis[ax].derivative > 1) high_order |= 1u << ax;
  }
  const int plane_bytes = ry * rs;

  // coordinates of up to 64 queries from position c of the segment travel global -> shared (the
  // lane's own two slots per axis: entries 2 * lane and 2 * lane + 1 of tq[ax]) while it computes
  auto prefetch = [&](const int64_t seg0, int c, int end) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int qi = c + 32 * k + lane;
      if (qi < end) {
#pragma unroll
        for (int ax = 0; ax < ND; ++ax)
          col_cp_async<(int)sizeof(T)>(xch_s + (unsigned)((ax * 64 + k) * (int)sizeof(T)), a.ax[ax].q + seg0 + qi);
      }
    }
  };

  bool prefetched = false;
  for (int k = 0;; ++k) {
    const int b = k & 1;
    st_mbar_wait(mbar_s + 8 * b, (unsigned)((k >> 1) & 1));
    const int4 hd = *reinterpret_cast<const int4*>(&HDR[b]);
    const int ws = hd.x, we = hd.y;
    if (ws >= we) break;
    const int ix0 = hd.z, iy0 = hd.w, ncol = HDR[b].ncol;
    const int64_t seg0 = ((int64_t)HDR[b].seg * gridDim.x + blockIdx.x) * seg;
    const int sub = (we - ws + kColWarps - 1) / kColWarps;
    int cur = min(ws + warp * sub, we);
    const int r1 = min(cur + sub, we);
    if (!prefetched && cur < r1) prefetch(seg0, cur, r1);
    prefetched = false;
    const char* const rows = rows0 + b * buf_bytes;
    T* const outs = a.out + seg0;
    const ColCell<T> cell_x = CELL[0][ix0];
    const unsigned flags0 = (high_order << 6) | (((unsigned)sign_word(cell_x.r) >> 31) << 6);

    // One round.  FAST: the round as the hot loop runs it, free of calls (out-of-line calls in
    // rare branches make the compiler spill around them in the loop); it gives up -- before
    // anything is written -- if a query of the round needs the one-query path or the general
    // taps, and the round is then redone by the other instantiation.
    auto round = [&](auto fast_tag) -> bool {
      constexpr bool FAST = decltype(fast_tag)::value;
      // ---- phase 1: two queries per lane (cur + lane, cur + 32 + lane) located ----
      st_cp_async_wait();
      T c[ND][2];
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) col_lds_pair(tq + ax * 64 + 2 * lane, c[ax][0], c[ax][1]);
      unsigned key[2], flg[2];  // flg: bits 0-5 fill, bits 6-8 axes that need the general taps
      T tt[2][ND];
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        bool slow = false;
        unsigned fl = flags0;
        int ii[ND];
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) {
          const int nk = a.ax[ax].nk;
          T v = c[ax][kk];
          if (IPXC_WRAP(ax)) {
            bool wrap_bad;
            v = wrap_one_period(v, T(fabs(P.axis[ax].period)), wrap_bad);
            slow = slow || wrap_bad;
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
            const T h = S.ks[ax][i];
            if (!((v >= cl.l || i == 1) && (v < h || i == nk - 1))) slow = true;
            if (can_fill[ax]) fl |= st_fill_bits(P.axis[ax], AXC[ax], i, nk, v) << (2 * ax);
          }
          tt[kk][ax] = t;
          ii[ax] = i;
        }
        slow = slow || (unsigned)(ii[1] - iy0) >= (unsigned)ncol;
        const bool valid = cur + 32 * kk + lane < r1;
        key[kk] = valid ? (((unsigned)ii[1] << 10) | (unsigned)ii[2] | (slow ? 1u << 30 : 0u)) : 0xffffffffu;
        flg[kk] = fl;
      }
      if constexpr (FAST) {
        const bool cold = (key[0] != 0xffffffffu && (((key[0] >> 30) & 1u) || (flg[0] & 0x1c0u))) ||
                          (key[1] != 0xffffffffu && (((key[1] >> 30) & 1u) || (flg[1] & 0x1c0u)));
        if (__any_sync(0xffffffffu, cold)) return false;  // (the coordinates are still in the buffer)
      }
      __syncwarp();  // everybody holds its coordinates: the buffer now takes the exchange
      // ---- re-pair: runs of equal keys; positions 2m, 2m+1 of a run share a lane ----
      const unsigned pk0 = __shfl_up_sync(0xffffffffu, key[0], 1);
      const unsigned pk1 = __shfl_up_sync(0xffffffffu, key[1], 1);
      const unsigned k31 = __shfl_sync(0xffffffffu, key[0], 31);
      const bool s0 = lane == 0 || key[0] != pk0;
      const bool s1 = key[1] != (lane == 0 ? k31 : pk1);
      const unsigned mlo = __ballot_sync(0xffffffffu, s0), mhi = __ballot_sync(0xffffffffu, s1);
      const int p0 = lane - (31 - __clz(mlo & le_mask));
      const unsigned mh = mhi & le_mask;
      const int p1 = 32 + lane - (mh ? 63 - __clz(mh) : 31 - __clz(mlo));
      const bool v0 = key[0] != 0xffffffffu, v1 = key[1] != 0xffffffffu;
      const unsigned olo = __ballot_sync(0xffffffffu, v0 && !(p0 & 1));
      const unsigned ohi = __ballot_sync(0xffffffffu, v1 && !(p1 & 1));
      const int nlo = __popc(olo);
      const int slot0 = __popc(olo & le_mask) - 1;
      const int slot1 = nlo + __popc(ohi & le_mask) - 1;
      const bool take0 = v0, take1 = v1 && slot1 < 32;  // (slot0 <= 31 always)
      // queries taken: all valid ones of the first half, and those of the second up to slot 31
      const int taken = __popc(__ballot_sync(0xffffffffu, v0)) + __popc(__ballot_sync(0xffffffffu, take1));
      const int nslots = min(32, nlo + __popc(ohi));
      // the entry after mine continues my run (only read for first halves)
      const bool cont0 = lane < 31 ? !((mlo >> (lane + 1)) & 1u) : !(mhi & 1u);
      const bool cont1 = lane < 31 && !((mhi >> (lane + 1)) & 1u);
      if (take0) {
        const int e = 2 * slot0 + (p0 & 1);
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) tq[ax * 64 + e] = tt[0][ax];
        meta[e] = make_uint2(key[0] | (cont0 ? 1u << 31 : 0u), (unsigned)lane | (flg[0] << 8));
      }
      if (take1) {
        const int e = 2 * slot1 + (p1 & 1);
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) tq[ax * 64 + e] = tt[1][ax];
        meta[e] = make_uint2(key[1] | (cont1 ? 1u << 31 : 0u), (unsigned)(32 + lane) | (flg[1] << 8));
      }
      __syncwarp();
      // ---- phase 2: the lane's pair ----
      const uint4 mm = *reinterpret_cast<const uint4*>(meta + 2 * lane);
      T ta[ND], tb[ND];
#pragma unroll
      for (int ax = 0; ax < ND; ++ax) col_lds_pair(tq + ax * 64 + 2 * lane, ta[ax], tb[ax]);
      __syncwarp();  // exchange read: the buffer takes the next round's coordinates
      const int nxt = cur + taken;
      if (nxt < r1) {
        prefetch(seg0, nxt, r1);
      } else {
        // last round of this window: start on the next one if its header is already there
        const int nb = b ^ 1;
        if (col_mbar_test(mbar_s + 8 * nb, (unsigned)(((k + 1) >> 1) & 1))) {
          const int ns = HDR[nb].s, ne = HDR[nb].e;
          if (ns < ne) {
            const int nsub = (ne - ns + kColWarps - 1) / kColWarps;
            const int n0 = min(ns + warp * nsub, ne);
            const int n1 = min(n0 + nsub, ne);
            if (n0 < n1) {
              prefetch(((int64_t)HDR[nb].seg * gridDim.x + blockIdx.x) * seg, n0, n1);
              prefetched = true;
            }
          }
        }
      }
      // the slow queries of this round (rare), by the lane that located them
      if constexpr (!FAST) {
        const bool slow0 = take0 && ((key[0] >> 30) & 1u), slow1 = take1 && ((key[1] >> 30) & 1u);
        if (slow0 || slow1) {
          if (slow0) st_redo<T, ND, ORD0>(&a, &P, smem, AXC, seg0 + cur + lane, 1u);
          if (slow1) st_redo<T, ND, ORD0>(&a, &P, smem, AXC, seg0 + cur + 32 + lane, 1u);
        }
      }
      const bool act_a = lane < nslots && !((mm.x >> 30) & 1u);
      const bool act_b = act_a && (mm.x >> 31) != 0u;
      if (act_a) {
        const int ix = ix0, iy = (mm.x >> 10) & 1023, iz = mm.x & 1023;
        T TA[ND][4], TB[ND][4];
        const unsigned gen = (mm.y >> 14) & 7u;  // (the pair shares the cell, hence these flags)
#pragma unroll
        for (int ax = 0; ax < ND; ++ax) {
          const int i = ax == 0 ? ix : (ax == 1 ? iy : iz);
          if (!FAST && ((gen >> ax) & 1u)) {
            const T rabs = t_abs(S.rr[ax][i]);
            const int order = ORD0 ? 0 : P.axis[ax].derivative;
            const Taps4<T> ga = st_taps_general_t<T>(S.ks[ax], S.ac[ax], i, ta[ax], rabs, order);
            const Taps4<T> gb = st_taps_general_t<T>(S.ks[ax], S.ac[ax], i, tb[ax], rabs, order);
            TA[ax][0] = ga.t0; TA[ax][1] = ga.t1; TA[ax][2] = ga.t2; TA[ax][3] = ga.t3;
            TB[ax][0] = gb.t0; TB[ax][1] = gb.t1; TB[ax][2] = gb.t2; TB[ax][3] = gb.t3;
          } else {
            const T r = ORD0 ? T(0) : S.rr[ax][i];
            col_taps<T, ORD0>(ta[ax], r, P.axis[ax].derivative, TA[ax]);
            col_taps<T, ORD0>(tb[ax], r, P.axis[ax].derivative, TB[ax]);
          }
        }
        const char* base = rows + (iy - iy0) * rs + (iz - 1) * (int)sizeof(T);
        T acc_a = T(0), acc_b = T(0);
#pragma unroll
        for (int aa = 0; aa < 4; ++aa) {
          T sya = T(0), syb = T(0);
#pragma unroll
          for (int bb = 0; bb < 4; ++bb) {
            const T* p = reinterpret_cast<const T*>(base + bb * rs);
            const T f0 = p[0], f1 = p[1], f2 = p[2], f3 = p[3];
            T ra = TA[2][0] * f0, rb = TB[2][0] * f0;
            ra = fma(TA[2][1], f1, ra); rb = fma(TB[2][1], f1, rb);
            ra = fma(TA[2][2], f2, ra); rb = fma(TB[2][2], f2, rb);
            ra = fma(TA[2][3], f3, ra); rb = fma(TB[2][3], f3, rb);
            sya = fma(TA[1][bb], ra, sya);
            syb = fma(TB[1][bb], rb, syb);
          }
          acc_a = fma(TA[0][aa], sya, acc_a);
          acc_b = fma(TB[0][aa], syb, acc_b);
          base += plane_bytes;
        }
        const unsigned fa = (mm.y >> 8) & 63u, fb = (mm.w >> 8) & 63u;
        __stcs(outs + (cur + (int)(mm.y & 63u)), st_apply_fill<T, ND>(P, fa, acc_a));
        if (act_b) __stcs(outs + (cur + (int)(mm.w & 63u)), st_apply_fill<T, ND>(P, fb, acc_b));
      }
      cur = nxt;
      return true;
    };
    while (cur < r1) {
      while (cur < r1 && round(std::true_type{})) {
      }
      if (cur < r1) round(std::false_type{});
    }
    __syncwarp();
    if (lane == 0) col_mbar_arrive(mbar_s + 16 + 8 * b);
  }
#undef IPXC_WRAP
}

// launch; IPX_EUNSUPPORTED when the shape does not apply (the caller uses another one)
template <typename T>
static int st_launch_column(const StArgs<T, 3>& a, const StDevice& d, cudaStream_t s) {
  const StLayout<3> lay = st_layout<T, 3>(a);
  long long cells = 0;
  for (int ax = 0; ax < 3; ++ax) {
    if (a.ax[ax].nk > 1023) return IPX_EUNSUPPORTED;  // cell indices travel in 10 bits each
    cells += ((long long)a.ax[ax].nk * (long long)sizeof(ColCell<T>) + 15) & ~15ll;
  }
  if (a.idx_out != nullptr) return IPX_EUNSUPPORTED;
  const long long row_bytes = (long long)a.pitch * (long long)sizeof(T);
  const long long rs = row_bytes;  // row stride in shared memory
  const long long fixed = (long long)lay.total + cells + (long long)kColWarps * ColCfg<T>::XCH_BYTES;
  const long long room = (long long)d.smem_optin - 2048 - fixed;
  long long ry = room / (2 * 4 * rs);
  if (ry > 32) ry = 32;
  if (ry < 5) return IPX_EUNSUPPORTED;  // fewer than two columns per window
  const size_t smem = (size_t)(fixed + 2 * 4 * ry * rs);
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
