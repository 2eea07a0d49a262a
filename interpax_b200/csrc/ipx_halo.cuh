// ipx_halo.cuh -- which table rows a row of the halo table stands for (ipx_halo_table_*), shared by
// the halo builder (ipx_stencil.cu) and the gradient kernels (ipx_grad.cu).
#pragma once
#include <stdint.h>

namespace ipx {

// Source rows of halo row u along one axis: one row with weight 1, or -- one row beyond a
// one-sided end -- the GHOST 2 v[end] - v[end -+ 1], with which the centred stencil
// (v[+1] - v[-1]) / 2h at the end knot equals the one-sided stencil (v[1] - v[0]) / h of
// _cubic1 / _cardinal (_fd_derivs.py:84-99, 314-320).
struct HaloSrc {
  int idx[2];
  int cnt;
};
__device__ __forceinline__ HaloSrc halo_src(int u, int n, int kind, const int32_t* __restrict__ perm) {
  HaloSrc h;
  h.cnt = 1;
  if (kind == 0) {
    const int p = u - 1;  // table row
    if (p < 0) { h.idx[0] = 0; h.idx[1] = 1; h.cnt = 2; }
    else if (p >= n) { h.idx[0] = n - 1; h.idx[1] = n - 2; h.cnt = 2; }
    else h.idx[0] = p;
  } else {
    auto wrapped = [&](int p) {  // padded position p in [0, n+1] -> table row
      int s = (p - 1) % n;
      if (s < 0) s += n;
      return perm != nullptr ? perm[s] : s;
    };
    const int p = u - 1;  // padded position, -1 .. n+2
    if (kind == 2 && p < 0) { h.idx[0] = wrapped(0); h.idx[1] = wrapped(1); h.cnt = 2; }
    else if (kind == 2 && p > n + 1) { h.idx[0] = wrapped(n + 1); h.idx[1] = wrapped(n); h.cnt = 2; }
    else h.idx[0] = wrapped(p);
  }
  return h;
}

}  // namespace ipx
