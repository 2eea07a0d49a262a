// ipx_prep.cu -- table/knot preparation around the evaluation kernels.
//
//   ipx_periodic_knots_*  knot side of _make_periodic (interpax/_spline.py:1117-1122):
//                         wrap, stable argsort, pad.  The table side (:1123-1135) is NOT
//                         materialised on the class path -- the evaluation kernel maps
//                         padded positions through `perm` instead.
//   ipx_pad_periodic_*    the materialised wrap-pad, only for the functional form that
//                         differentiates the padded table (_spline.py:924-938 + 1027-1040).
//   ipx_pack_*            SoA -> AoS interleave of the 2^d tables (layout in DESIGN.md).
#include <cuda_runtime.h>

#include "ipx_device.cuh"
#include "ipx_internal.cuh"

namespace ipx {

constexpr int kPrepThreads = 256;

static int prep_grid(int64_t total) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess)
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t need = (total + kPrepThreads - 1) / kPrepThreads;
  int64_t cap = (int64_t)sms * 16;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

// workspace layout: xm[0..n) wrapped knots, then two int32 counters {descents, last descent}
template <typename T>
__global__ void wrap_knots_kernel(const T* __restrict__ x, int n, T P, T* __restrict__ xm,
                                  int* __restrict__ counters) {
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    T v = py_mod(x[k], P);
    xm[k] = v;
    if (k + 1 < n) {
      T nx = py_mod(x[k + 1], P);
      if (v > nx) {
        atomicAdd(&counters[0], 1);
        atomicMax(&counters[1], k);
      }
    }
  }
}

__global__ void zero_counters_kernel(int* counters) { counters[0] = 0; counters[1] = -1; }

// stable rank of every wrapped knot.
//   0 descents: already sorted.
//   1 descent (the usual case: knots spanning at most one period): two sorted runs
//     A = [0..d], B = [d+1..n); stable merge by binary search (ties: A first).
//   otherwise: O(n^2) counting rank (correct for anything; only pathological input gets here).
template <typename T>
__global__ void rank_knots_kernel(const T* __restrict__ xm, int n, T P,
                                  const int* __restrict__ counters, T* __restrict__ xpad,
                                  int32_t* __restrict__ perm) {
  const int descents = counters[0];
  const int d = counters[1];
  const int stride = gridDim.x * blockDim.x;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) {
    const T v = xm[k];
    int r;
    if (descents == 0) {
      r = k;
    } else if (descents == 1) {
      if (k <= d) {
        // #B strictly less than v
        int lo = d + 1, hi = n;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (xm[mid] < v) lo = mid + 1; else hi = mid; }
        r = k + (lo - (d + 1));
      } else {
        // #A less than or equal to v
        int lo = 0, hi = d + 1;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (xm[mid] <= v) lo = mid + 1; else hi = mid; }
        r = (k - (d + 1)) + lo;
      }
    } else {
      r = 0;
      for (int j = 0; j < n; ++j) {
        T u = xm[j];
        r += (u < v) || (u == v && j < k);
      }
    }
    perm[r] = k;
    xpad[r + 1] = v;
    if (r == n - 1) xpad[0] = v - P;
    if (r == 0) xpad[n + 1] = v + P;
  }
}

template <typename T>
__global__ void pad_periodic_kernel(const T* __restrict__ in, int64_t outer, int n, int64_t inner,
                                    const int32_t* __restrict__ perm, T* __restrict__ out) {
  const int64_t total = outer * (n + 2) * inner;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    int64_t i = e % inner;
    int64_t r = e / inner;
    int p = (int)(r % (n + 2));
    int64_t o = r / (n + 2);
    int s = p - 1;
    if (s < 0) s += n;
    if (s >= n) s -= n;
    int src = perm ? perm[s] : s;
    out[e] = in[(o * n + src) * inner + i];
  }
}

template <typename T, int NT> struct PackArgs {
  const T* __restrict__ tab[NT];  // canonical record order
};

template <typename T, int NT>
__global__ void pack_kernel(const PackArgs<T, NT> a, int64_t total, T* __restrict__ out) {
  // one thread per output element: coalesced writes, NT interleaved coalesced read streams
  const int64_t n_out = total * NT;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_out; e += stride) {
    int slot = (int)(e % NT);
    int64_t rec = e / NT;
    const T* src = a.tab[0];
#pragma unroll
    for (int k = 1; k < NT; ++k) if (slot == k) src = a.tab[k];
    out[e] = src[rec];
  }
}

template <typename T>
static int launch_periodic_knots(const T* x, int32_t n, double period, T* xpad, int32_t* perm,
                                 void* workspace, void* stream) {
  if (x == nullptr || xpad == nullptr || perm == nullptr || workspace == nullptr || n < 1)
    return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  T* xm = static_cast<T*>(workspace);
  int* counters = reinterpret_cast<int*>(xm + n + ((n & 1) ? 1 : 0));  // keep 8-byte alignment
  T P = T(period < 0 ? -period : period);
  zero_counters_kernel<<<1, 1, 0, s>>>(counters); count_launches(1);
  wrap_knots_kernel<T><<<prep_grid(n), kPrepThreads, 0, s>>>(x, n, P, xm, counters); count_launches(1);
  rank_knots_kernel<T><<<prep_grid(n), kPrepThreads, 0, s>>>(xm, n, P, counters, xpad, perm); count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_pad(const T* in, int64_t outer, int32_t n, int64_t inner, const int32_t* perm,
                      T* out, void* stream) {
  if (outer < 0 || inner < 0 || n < 1) return IPX_EINVAL;
  if (outer * inner == 0) return IPX_OK;
  if (in == nullptr || out == nullptr) return IPX_EINVAL;
  pad_periodic_kernel<T><<<prep_grid(outer * (n + 2) * inner), kPrepThreads, 0,
                           static_cast<cudaStream_t>(stream)>>>(in, outer, n, inner, perm, out); count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

// reference stacking order -> canonical record slot (same rule as in ipx_eval.cu)
static int pack_slot(int ndim, int ff) {
  static const int d1[2] = {0, 1};
  static const int d2[4] = {0, 1, 2, 3};
  static const int d3[8] = {0, 1, 2, 4, 3, 5, 6, 7};
  int mask = ndim == 1 ? d1[ff] : ndim == 2 ? d2[ff] : d3[ff];
  int slot = 0;
  for (int ax = 0; ax < ndim; ++ax)
    if (mask & (1 << ax)) slot |= 1 << (ndim - 1 - ax);
  return slot;
}

template <typename T, int NT>
static int launch_pack_n(int ndim, const T* const* tables, int64_t total, T* out, cudaStream_t s) {
  PackArgs<T, NT> a;
  for (int ff = 0; ff < NT; ++ff) {
    if (tables[ff] == nullptr) return IPX_EINVAL;
    a.tab[pack_slot(ndim, ff)] = tables[ff];
  }
  pack_kernel<T, NT><<<prep_grid(total * NT), kPrepThreads, 0, s>>>(a, total, out); count_launches(1);
  return cudaPeekAtLastError() == cudaSuccess ? IPX_OK : IPX_ELAUNCH;
}

template <typename T>
static int launch_pack(int32_t ndim, const T* const* tables, int64_t total, T* out, void* stream) {
  if (tables == nullptr || total < 0 || ndim < 1 || ndim > 3) return IPX_EINVAL;
  if (total == 0) return IPX_OK;
  if (out == nullptr) return IPX_EINVAL;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ndim == 1) return launch_pack_n<T, 2>(ndim, tables, total, out, s);
  if (ndim == 2) return launch_pack_n<T, 4>(ndim, tables, total, out, s);
  return launch_pack_n<T, 8>(ndim, tables, total, out, s);
}

}  // namespace ipx

extern "C" int ipx_periodic_knots_f32(const float* x, int32_t n, double period, float* xpad,
                                      int32_t* perm, void* workspace, void* stream) {
  return ipx::launch_periodic_knots<float>(x, n, period, xpad, perm, workspace, stream);
}
extern "C" int ipx_periodic_knots_f64(const double* x, int32_t n, double period, double* xpad,
                                      int32_t* perm, void* workspace, void* stream) {
  return ipx::launch_periodic_knots<double>(x, n, period, xpad, perm, workspace, stream);
}
extern "C" int ipx_pad_periodic_f32(const float* in, int64_t outer, int32_t n, int64_t inner,
                                    const int32_t* perm, float* out, void* stream) {
  return ipx::launch_pad<float>(in, outer, n, inner, perm, out, stream);
}
extern "C" int ipx_pad_periodic_f64(const double* in, int64_t outer, int32_t n, int64_t inner,
                                    const int32_t* perm, double* out, void* stream) {
  return ipx::launch_pad<double>(in, outer, n, inner, perm, out, stream);
}
extern "C" int ipx_pack_f32(int32_t ndim, const float* const* tables, int64_t nodes_times_batch,
                            float* out, void* stream) {
  return ipx::launch_pack<float>(ndim, tables, nodes_times_batch, out, stream);
}
extern "C" int ipx_pack_f64(int32_t ndim, const double* const* tables, int64_t nodes_times_batch,
                            double* out, void* stream) {
  return ipx::launch_pack<double>(ndim, tables, nodes_times_batch, out, stream);
}
