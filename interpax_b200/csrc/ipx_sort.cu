// ipx_sort.cu -- spatial pre-sort ("binning") of the query stream.
//
// Why: the cubic kernels are fast when neighbouring queries fall into the same or adjacent
// table cells (ipx_eval_tile.cuh) and slow when every query pulls its own 512 B of table rows
// out of DRAM.  A caller whose queries arrive in random order can ask for them to be visited
// in cell order instead:
//     order = argsort(cell key)           ipx_cell_order_*   (this file: keys + LSD radix sort)
//     qs    = q[order]                    ipx_gather_*
//     out_s = eval(qs)                    ipx_eval*d_*       (unchanged)
//     out[order] = out_s                  ipx_scatter_*
// Results are bit-identical to the unsorted call because every query is evaluated by the same
// arithmetic wherever it sits in the stream.
//
// This has no counterpart in the reference (interpax evaluates queries in the order given,
// _spline.py:1051-1099); it is SURVEY section 8(f) row f4.  The cell key is built from the same
// interval lookup as the evaluation (wrap + searchsorted-right + clip), so "same key" means
// "same cell" exactly.
//
// Radix sort: least-significant-digit, 8 bits per pass, only as many passes as the key has
// bits.  Per pass: per-block digit histogram -> exclusive scan over (digit, block) -> stable
// scatter (ranks inside a warp from __match_any_sync, warps and blocks ordered by the scan).
#include <cuda_runtime.h>

#include "ipx_device.cuh"
#include "ipx_internal.cuh"

namespace ipx {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;                           // keys per thread per tile
constexpr int kSortTile = kSortThreads * kSortItems;     // 4096 keys per block
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

static int sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// ---------------------------------------------------------------------------------- keys
template <typename T, int ND> struct KeyArgs {
  AxisDesc<T> ax[ND];
  uint32_t stride[ND];  // key = sum cell_ax * stride[ax]; last axis has stride 1
  uint32_t cells[ND];   // cells along the axis; on a periodic axis the padded intervals 1 and
                        // n+1 are the same physical cell (the seam) and share a key digit
  int64_t nq;
  ipx_params hp;
  const ipx_params* dp;
};

// Interleaved copy of the query coordinates, one record per query padded to a power of two
// (3-D: x, y, z, 0), so that the random gather after the sort touches ONE 32-byte sector per
// query instead of one per coordinate array.
template <int ND> struct CoordRec { static constexpr int kLen = ND == 3 ? 4 : ND; };

template <typename T, int ND>
__global__ void cell_key_kernel(const KeyArgs<T, ND> a, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals, T* __restrict__ recs) {
  __shared__ AxisConst<T> consts[ND];
  __shared__ ipx_params params_s;
  if (threadIdx.x < ND) axis_consts(a.ax[threadIdx.x], consts[threadIdx.x]);
  if (threadIdx.x == 32) params_s = a.dp ? *a.dp : a.hp;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // (the coordinates of the thread's NEXT element are requested before the current one is located:
  // the kernel waits on these loads, 66 % of its stall samples)
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  T qn[ND];
#pragma unroll
  for (int ax = 0; ax < ND; ++ax) qn[ax] = e < a.nq ? __ldcs(a.ax[ax].q + e) : T(0);
  for (; e < a.nq; e += stride) {
    T qc[ND];
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) qc[ax] = qn[ax];
    const int64_t en = e + stride;
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) qn[ax] = en < a.nq ? __ldcs(a.ax[ax].q + en) : T(0);
    uint32_t key = 0;
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) {
      AxisLoc<T> L;
      locate_value(a.ax[ax], consts[ax], params_s.axis[ax], qc[ax], L);
      uint32_t c = (uint32_t)(L.i - 1);
      if (c >= a.cells[ax]) c -= a.cells[ax];
      key += c * a.stride[ax];
    }
    keys[e] = key;
    vals[e] = (uint32_t)e;
    if (recs != nullptr) {
      constexpr int R = CoordRec<ND>::kLen;
      T c[R];
#pragma unroll
      for (int ax = 0; ax < R; ++ax) c[ax] = ax < ND ? qc[ax < ND ? ax : 0] : T(0);
      T* r = recs + e * R;
      if constexpr (R == 1) {
        r[0] = c[0];
      } else if constexpr (sizeof(T) == 8) {
#pragma unroll
        for (int m = 0; m < R; m += 2) *reinterpret_cast<double2*>(r + m) = make_double2(c[m], c[m + 1]);
      } else if constexpr (R == 2) {
        *reinterpret_cast<float2*>(r) = make_float2(c[0], c[1]);
      } else {
        *reinterpret_cast<float4*>(r) = make_float4(c[0], c[1], c[2], c[3]);
      }
    }
  }
}

// dst[ax][k] = recs[order[k]][ax]: sequential writes, one random record read per query
template <typename T, int ND> struct RecGatherArgs {
  T* __restrict__ dst[ND];
};

template <typename T, int ND>
__global__ void gather_records_kernel(const T* __restrict__ recs, const uint32_t* __restrict__ order,
                                      int64_t n, const RecGatherArgs<T, ND> a) {
  constexpr int R = CoordRec<ND>::kLen;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const T* r = recs + (int64_t)order[k] * R;
    T c[R];
    if constexpr (R == 1) {
      c[0] = __ldg(r);
    } else if constexpr (sizeof(T) == 8) {
#pragma unroll
      for (int m = 0; m < R; m += 2) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(r + m));
        c[m] = v.x; c[m + 1] = v.y;
      }
    } else if constexpr (R == 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(r));
      c[0] = v.x; c[1] = v.y;
    } else {
      const float4 v = __ldg(reinterpret_cast<const float4*>(r));
      c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
    }
#pragma unroll
    for (int ax = 0; ax < ND; ++ax) a.dst[ax][k] = c[ax];
  }
}

// ---------------------------------------------------------------------------------- radix passes
__global__ void radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift,
                                  uint32_t* __restrict__ hist, int nblocks) {
  __shared__ uint32_t h[kRadix];
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) h[d] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
  for (int it = 0; it < kSortItems; ++it) {
    const int64_t e = base + it * kSortThreads + threadIdx.x;
    if (e < n) atomicAdd(&h[(keys[e] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += kSortThreads) hist[(int64_t)d * nblocks + blockIdx.x] = h[d];
}

// exclusive scan of `count` uint32 values in three steps (tiles of kScanTile per block)
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  __shared__ uint32_t block_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    uint32_t o = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += o;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
    uint32_t winc = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, winc, off);
      if (lane >= off) winc += o;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = winc - w;
    if (lane == kScanThreads / 32 - 1) block_total = winc;
  }
  __syncthreads();
  uint32_t excl = inc - v + warp_sums[warp];
  if (total) *total = block_total;
  __syncthreads();
  return excl;
}

__global__ void scan_tiles_kernel(uint32_t* __restrict__ data, int64_t count, uint32_t* __restrict__ tile_sums) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = base + k < count ? data[base + k] : 0;
    s += v[k];
  }
  uint32_t total;
  uint32_t excl = block_exclusive_scan(s, &total);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < count) data[base + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_sums_kernel(uint32_t* __restrict__ sums, int ntiles) {
  // single block: sequential over chunks of kScanThreads, carry in a register
  uint32_t carry = 0;
  for (int base = 0; base < ntiles; base += kScanThreads) {
    const int i = base + threadIdx.x;
    uint32_t v = i < ntiles ? sums[i] : 0;
    uint32_t total;
    uint32_t excl = block_exclusive_scan(v, &total);
    if (i < ntiles) sums[i] = excl + carry;
    carry += total;
  }
}

__global__ void scan_add_kernel(uint32_t* __restrict__ data, int64_t count, const uint32_t* __restrict__ sums) {
  const uint32_t add = sums[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  for (int k = threadIdx.x; k < kScanTile; k += kScanThreads)
    if (base + k < count) data[base + k] += add;
}

// stable scatter of one tile per block.  Warp w owns items [w*512, (w+1)*512) of the tile and
// walks them 32 at a time in order, so rank-within-warp preserves the input order.  Items are
// first placed in shared memory in digit order, then written out: consecutive threads write
// consecutive addresses of a digit's run instead of 32 scattered words per warp.
#ifndef IPX_SORT_MINBLOCKS
#define IPX_SORT_MINBLOCKS 4  // 64 registers: four blocks per SM (the kernel waits on its loads: 1.21 -> 0.9 ms per pass of 1e8 pairs)
#endif
__global__ void __launch_bounds__(kSortThreads, IPX_SORT_MINBLOCKS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, int64_t n,
                     int shift, const uint32_t* __restrict__ offsets, int nblocks,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  constexpr int kWarps = kSortThreads / 32;
  constexpr int kPerWarp = kSortTile / kWarps;  // 512
  constexpr int kSteps = kPerWarp / 32;         // 16
  static_assert(kRadix == kSortThreads, "one thread per digit in the prefix step");
  __shared__ uint32_t wcount[kWarps][kRadix];   // per-warp digit counts, then warp-exclusive bases
  __shared__ uint32_t dbase[kRadix];            // block-local exclusive base of each digit
  __shared__ uint32_t gbase[kRadix];            // global position of the digit's run minus dbase
  __shared__ uint32_t skeys[kSortTile];
  __shared__ uint32_t svals[kSortTile];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int d = lane; d < kRadix; d += 32) wcount[warp][d] = 0;
  __syncwarp();
  const int64_t tbase = (int64_t)blockIdx.x * kSortTile;
  const int64_t wbase = tbase + (int64_t)warp * kPerWarp;
  uint32_t key[kSteps], val[kSteps];
  uint16_t rank[kSteps];
#pragma unroll
  for (int s = 0; s < kSteps; ++s) {
    const int64_t e = wbase + s * 32 + lane;
    const bool ok = e < n;
    key[s] = ok ? keys_in[e] : 0xffffffffu;
    val[s] = ok ? vals_in[e] : 0;
    const unsigned d = ok ? (key[s] >> shift) & (kRadix - 1) : kRadix;  // kRadix = "no item"
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const unsigned before = __popc(peers & ((1u << lane) - 1));
    uint32_t basecnt = 0;
    if (ok) basecnt = wcount[warp][d];  // all peers read the same value before the leader bumps it
    __syncwarp();
    if (ok && before == 0) wcount[warp][d] = basecnt + __popc(peers);
    __syncwarp();
    rank[s] = (uint16_t)(basecnt + before);
  }
  __syncthreads();
  // one thread per digit: warp-exclusive bases, digit total, block-exclusive digit base
  {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t c = wcount[w][d];
      wcount[w][d] = run;
      run += c;
    }
    const uint32_t excl = block_exclusive_scan(run, nullptr);
    dbase[d] = excl;
    gbase[d] = offsets[(int64_t)d * nblocks + blockIdx.x] - excl;
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kSteps; ++s) {
    const int64_t e = wbase + s * 32 + lane;
    if (e < n) {
      const unsigned d = (key[s] >> shift) & (kRadix - 1);
      const uint32_t local = dbase[d] + wcount[warp][d] + rank[s];
      skeys[local] = key[s];
      svals[local] = val[s];
    }
  }
  __syncthreads();
  const int count = (int)min((int64_t)kSortTile, n - tbase);
  for (int i = threadIdx.x; i < count; i += kSortThreads) {
    const uint32_t k = skeys[i];
    const uint32_t pos = gbase[(k >> shift) & (kRadix - 1)] + (uint32_t)i;
    keys_out[pos] = k;
    vals_out[pos] = svals[i];
  }
}

// workspace layout (all uint32): keys_a[n] keys_b[n] vals_b[n] hist[kRadix*nblocks] sums[ntiles]
struct SortPlan {
  int64_t n;
  int nblocks;
  int64_t hist_count;
  int ntiles;
  size_t off_keys_a, off_keys_b, off_vals_b, off_hist, off_sums, total;
};

static SortPlan make_plan(int64_t n) {
  SortPlan p;
  p.n = n;
  p.nblocks = (int)((n + kSortTile - 1) / kSortTile);
  if (p.nblocks < 1) p.nblocks = 1;
  p.hist_count = (int64_t)kRadix * p.nblocks;
  p.ntiles = (int)((p.hist_count + kScanTile - 1) / kScanTile);
  auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
  size_t off = 0;
  p.off_keys_a = off; off = align(off + (size_t)n * 4);
  p.off_keys_b = off; off = align(off + (size_t)n * 4);
  p.off_vals_b = off; off = align(off + (size_t)n * 4);
  p.off_hist = off; off = align(off + (size_t)p.hist_count * 4);
  p.off_sums = off; off = align(off + (size_t)p.ntiles * 4);
  p.total = off;
  return p;
}

// sorts (keys_a, vals_a) by the low `bits` bits of the key; the sorted values end in vals_a
static void radix_sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b,
                             uint32_t* hist, uint32_t* sums, const SortPlan& p, int bits, cudaStream_t s) {
  int passes = (bits + kRadixBits - 1) / kRadixBits;
  if (passes < 1) passes = 1;
  uint32_t *ki = keys_a, *vi = vals_a, *ko = keys_b, *vo = vals_b;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = pass * kRadixBits;
    radix_hist_kernel<<<p.nblocks, kSortThreads, 0, s>>>(ki, p.n, shift, hist, p.nblocks); count_launches(1);
    scan_tiles_kernel<<<p.ntiles, kScanThreads, 0, s>>>(hist, p.hist_count, sums); count_launches(1);
    scan_sums_kernel<<<1, kScanThreads, 0, s>>>(sums, p.ntiles); count_launches(1);
    scan_add_kernel<<<p.ntiles, kScanThreads, 0, s>>>(hist, p.hist_count, sums); count_launches(1);
    radix_scatter_kernel<<<p.nblocks, kSortThreads, 0, s>>>(ki, vi, p.n, shift, hist, p.nblocks, ko, vo); count_launches(1);
    uint32_t* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  if (vi != vals_a)  // odd number of passes: the permutation sits in the (b) buffer
    cudaMemcpyAsync(vals_a, vi, (size_t)p.n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s);
}

template <typename T, int ND>
static int launch_cell_order(const T* const* q, int64_t nq, const T* const* knots, const int32_t* n,
                             int32_t periodic_mask, const ipx_params* params, int32_t params_on_device,
                             uint32_t* order, T* const* q_sorted, void* workspace, cudaStream_t s,
                             T* recs_out = nullptr) {
  KeyArgs<T, ND> a;
  double cells = 1.0;
  for (int ax = 0; ax < ND; ++ax) {
    if (q[ax] == nullptr || knots[ax] == nullptr || n[ax] < 2) return IPX_EINVAL;
    const int per = (periodic_mask >> ax) & 1;
    const int wrap_only = (periodic_mask >> (4 + ax)) & 1;
    a.ax[ax].q = q[ax];
    a.ax[ax].knots = knots[ax];
    a.ax[ax].perm = nullptr;  // keys need the interval index only
    a.ax[ax].n = n[ax];
    a.ax[ax].nk = per ? n[ax] + 2 : n[ax];
    a.ax[ax].periodic = 0;
    a.ax[ax].wrap = per | wrap_only;
    a.cells[ax] = (uint32_t)(per ? n[ax] : n[ax] - 1);
    cells *= (double)a.cells[ax];
  }
  if (cells >= 4294967296.0) return IPX_EUNSUPPORTED;
  uint32_t stride = 1;
  for (int ax = ND - 1; ax >= 0; --ax) {
    a.stride[ax] = stride;
    stride *= a.cells[ax];
  }
  int bits = 1;
  while (bits < 32 && (double)(1ull << bits) < cells) ++bits;
  a.nq = nq;
  if (params_on_device) { a.dp = params; a.hp = ipx_params{}; } else { a.dp = nullptr; a.hp = *params; }

  const SortPlan p = make_plan(nq);
  char* ws = static_cast<char*>(workspace);
  uint32_t* keys_a = reinterpret_cast<uint32_t*>(ws + p.off_keys_a);
  uint32_t* keys_b = reinterpret_cast<uint32_t*>(ws + p.off_keys_b);
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(ws + p.off_vals_b);
  uint32_t* hist = reinterpret_cast<uint32_t*>(ws + p.off_hist);
  uint32_t* sums = reinterpret_cast<uint32_t*>(ws + p.off_sums);
  int64_t need = (nq + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 16;
  T* recs = recs_out;  // the interleaved coordinate records, left to the caller (no gather)
  if (q_sorted != nullptr) {
    for (int ax = 0; ax < ND; ++ax)
      if (q_sorted[ax] == nullptr) return IPX_EINVAL;
    recs = reinterpret_cast<T*>(ws + p.total);  // p.total is 256-byte aligned
  }
  cell_key_kernel<T, ND><<<(int)(need < cap ? need : cap), 256, 0, s>>>(a, keys_a, order, recs); count_launches(1);
  radix_sort_pairs(keys_a, order, keys_b, vals_b, hist, sums, p, bits, s);
  if (q_sorted != nullptr) {
    RecGatherArgs<T, ND> g;
    for (int ax = 0; ax < ND; ++ax) g.dst[ax] = q_sorted[ax];
    int64_t gcap = (int64_t)sm_count() * 32;
    gather_records_kernel<T, ND><<<(int)(need < gcap ? need : gcap), 256, 0, s>>>(recs, order, nq, g); count_launches(1);
  }
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
__global__ void gather_kernel(const T* __restrict__ src, const uint32_t* __restrict__ order, int64_t n,
                              int64_t batch, T* __restrict__ dst) {
  const int64_t total = n * batch;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t k = batch == 1 ? e : e / batch;
    const int64_t b = e - k * batch;
    dst[e] = src[(int64_t)order[k] * batch + b];
  }
}

template <typename T>
__global__ void scatter_kernel(const T* __restrict__ src, const uint32_t* __restrict__ order, int64_t n,
                               int64_t batch, T* __restrict__ dst) {
  const int64_t total = n * batch;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int64_t k = batch == 1 ? e : e / batch;
    const int64_t b = e - k * batch;
    dst[(int64_t)order[k] * batch + b] = src[e];
  }
}

template <typename T, int NA> struct MultiArgs {
  const T* __restrict__ src[NA];
  T* __restrict__ dst[NA];
};

// dst[a][k] = src[a][order[k]] for NA arrays at once: the order is read once and the NA random
// loads of a thread are in flight together
template <typename T, int NA>
__global__ void gather_multi_kernel(const MultiArgs<T, NA> a, const uint32_t* __restrict__ order, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const uint32_t o = order[k];
    T v[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) v[i] = a.src[i][o];
#pragma unroll
    for (int i = 0; i < NA; ++i) a.dst[i][k] = v[i];
  }
}

template <typename T, int NA>
static int launch_gather_multi_n(const T* const* src, const uint32_t* order, int64_t n, T* const* dst,
                                 cudaStream_t s) {
  MultiArgs<T, NA> a;
  for (int i = 0; i < NA; ++i) {
    if (src[i] == nullptr || dst[i] == nullptr) return IPX_EINVAL;
    a.src[i] = src[i];
    a.dst[i] = dst[i];
  }
  int64_t need = (n + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 32;
  gather_multi_kernel<T, NA><<<(int)(need < cap ? need : cap), 256, 0, s>>>(a, order, n); count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_gather_multi(int32_t narrays, const T* const* src, const uint32_t* order, int64_t n,
                               T* const* dst, void* stream) {
  if (narrays < 1 || narrays > 3 || src == nullptr || dst == nullptr || n < 0) return IPX_EINVAL;
  if (n == 0) return IPX_OK;
  if (order == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (narrays == 1) return launch_gather_multi_n<T, 1>(src, order, n, dst, s);
  if (narrays == 2) return launch_gather_multi_n<T, 2>(src, order, n, dst, s);
  return launch_gather_multi_n<T, 3>(src, order, n, dst, s);
}

template <typename T, bool GATHER>
static int launch_permute(const T* src, const uint32_t* order, int64_t n, int64_t batch, T* dst, void* stream) {
  if (n < 0 || batch < 1) return IPX_EINVAL;
  if (n == 0) return IPX_OK;
  if (src == nullptr || order == nullptr || dst == nullptr) return IPX_EINVAL;
  int64_t need = (n * batch + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 32;
  const int grid = (int)(need < cap ? need : cap);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (GATHER) gather_kernel<T><<<grid, 256, 0, s>>>(src, order, n, batch, dst);
  else scatter_kernel<T><<<grid, 256, 0, s>>>(src, order, n, batch, dst);
  count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int cell_order_dispatch(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,
                               const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                               int32_t params_on_device, uint32_t* order, T* const* q_sorted,
                               void* workspace, void* stream, T* recs_out = nullptr) {
  if (ndim < 1 || ndim > 3 || q == nullptr || knots == nullptr || n == nullptr || params == nullptr || nq < 0)
    return IPX_EINVAL;
  if (nq >= 4294967296LL) return IPX_EUNSUPPORTED;
  if (nq == 0) return IPX_OK;
  if (order == nullptr || workspace == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ndim == 1) return launch_cell_order<T, 1>(q, nq, knots, n, periodic_mask, params, params_on_device, order, q_sorted, workspace, s, recs_out);
  if (ndim == 2) return launch_cell_order<T, 2>(q, nq, knots, n, periodic_mask, params, params_on_device, order, q_sorted, workspace, s, recs_out);
  return launch_cell_order<T, 3>(q, nq, knots, n, periodic_mask, params, params_on_device, order, q_sorted, workspace, s, recs_out);
}

}  // namespace ipx

extern "C" size_t ipx_cell_order_workspace_bytes(int64_t nq) {
  return ipx::make_plan(nq < 0 ? 0 : nq).total;
}
extern "C" int ipx_cell_order_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                                  const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                                  int32_t params_on_device, uint32_t* order, void* workspace, void* stream) {
  return ipx::cell_order_dispatch<float>(ndim, q, nq, knots, n, periodic_mask, params, params_on_device, order,
                                         nullptr, workspace, stream);
}
extern "C" int ipx_cell_sort_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                                 const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                                 int32_t params_on_device, uint32_t* order, float* const* q_sorted,
                                 void* workspace, void* stream) {
  if (q_sorted == nullptr) return IPX_EINVAL;
  return ipx::cell_order_dispatch<float>(ndim, q, nq, knots, n, periodic_mask, params, params_on_device, order,
                                         q_sorted, workspace, stream);
}
extern "C" int ipx_cell_order_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                                  const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                                  int32_t params_on_device, uint32_t* order, void* workspace, void* stream) {
  return ipx::cell_order_dispatch<double>(ndim, q, nq, knots, n, periodic_mask, params, params_on_device, order,
                                          nullptr, workspace, stream);
}
extern "C" int ipx_cell_sort_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                                 const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                                 int32_t params_on_device, uint32_t* order, double* const* q_sorted,
                                 void* workspace, void* stream) {
  if (q_sorted == nullptr) return IPX_EINVAL;
  return ipx::cell_order_dispatch<double>(ndim, q, nq, knots, n, periodic_mask, params, params_on_device, order,
                                          q_sorted, workspace, stream);
}
#define IPX_DEFINE_ORDER_RECORDS(SUF, T)                                                                          \
  extern "C" int ipx_cell_order_records_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots, \
                                              const int32_t* n, int32_t periodic_mask, const ipx_params* params,  \
                                              int32_t params_on_device, uint32_t* order, T* records,              \
                                              void* workspace, void* stream) {                                    \
    if (records == nullptr && nq > 0) return IPX_EINVAL;                                                          \
    return ipx::cell_order_dispatch<T>(ndim, q, nq, knots, n, periodic_mask, params, params_on_device, order,     \
                                       nullptr, workspace, stream, records);                                      \
  }
IPX_DEFINE_ORDER_RECORDS(f32, float)
IPX_DEFINE_ORDER_RECORDS(f64, double)

extern "C" size_t ipx_cell_sort_workspace_bytes(int32_t ndim, int64_t nq, int32_t elem_size) {
  if (nq < 0) nq = 0;
  const size_t rec = ndim == 3 ? 4 : (ndim < 1 ? 1 : (size_t)ndim);
  return ipx::make_plan(nq).total + (size_t)nq * rec * (size_t)(elem_size == 4 ? 4 : 8);
}
extern "C" int ipx_gather_f32(const float* src, const uint32_t* order, int64_t n, int64_t batch, float* dst,
                              void* stream) {
  return ipx::launch_permute<float, true>(src, order, n, batch, dst, stream);
}
extern "C" int ipx_gather_f64(const double* src, const uint32_t* order, int64_t n, int64_t batch, double* dst,
                              void* stream) {
  return ipx::launch_permute<double, true>(src, order, n, batch, dst, stream);
}
extern "C" int ipx_gather_multi_f32(int32_t narrays, const float* const* src, const uint32_t* order, int64_t n,
                                    float* const* dst, void* stream) {
  return ipx::launch_gather_multi<float>(narrays, src, order, n, dst, stream);
}
extern "C" int ipx_gather_multi_f64(int32_t narrays, const double* const* src, const uint32_t* order,
                                    int64_t n, double* const* dst, void* stream) {
  return ipx::launch_gather_multi<double>(narrays, src, order, n, dst, stream);
}
extern "C" int ipx_scatter_f32(const float* src, const uint32_t* order, int64_t n, int64_t batch, float* dst,
                               void* stream) {
  return ipx::launch_permute<float, false>(src, order, n, batch, dst, stream);
}
extern "C" int ipx_scatter_f64(const double* src, const uint32_t* order, int64_t n, int64_t batch, double* dst,
                               void* stream) {
  return ipx::launch_permute<double, false>(src, order, n, batch, dst, stream);
}
