/* ipx.h -- C ABI of libipx.so, the B200 (sm_100a) query-evaluation path for interpax.
 *
 * This is the drop-in boundary.  The reference (f0uriest/interpax) has no FFI seam of
 * its own -- the path is inline jax.numpy under jax.jit -- so each entry point below
 * names the reference code whose *body* it replaces (paths relative to the reference
 * repo).  A binding (jax.ffi custom call, ctypes, ...) marshals arrays into these
 * calls; see INTEGRATION.md.
 *
 * Conventions (all entry points):
 *   - every data pointer is a DEVICE pointer owned by the caller; outputs are
 *     pre-allocated by the caller; nothing is allocated, freed or synchronised here;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and nowhere else,
 *     so every call can be captured into a CUDA graph;
 *   - operands that are *traced* values in the reference (derivative orders, extrap
 *     flags and fill values, period values: only `method` is static under jit,
 *     interpax/_spline.py:444,595,808) travel in `ipx_params`, which may live on the
 *     host (copied into the launch) or on the device (read by the kernel);
 *   - the return value is IPX_OK or a negative IPX_E* code; nothing aborts or throws;
 *   - re-entrant and thread-safe: no mutable global state.
 */
#ifndef IPX_H_
#define IPX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPX_VERSION 100 /* 0.1.0 */

#define IPX_OK 0
#define IPX_EINVAL (-1)       /* bad argument (null pointer, negative size, bad enum) */
#define IPX_EUNSUPPORTED (-2) /* valid request this build does not implement */
#define IPX_ELAUNCH (-3)      /* cudaPeekAtLastError() reported a launch failure */

/* method class: which branch of interp1d/2d/3d runs (interpax/_spline.py:529,540,569).
 * All of cubic/cubic2/cardinal/catmull-rom/akima/monotonic/monotonic-0 are IPX_CUBIC:
 * they differ only in how the slope tables were produced (ipx_slopes_*). */
#define IPX_NEAREST 0
#define IPX_LINEAR 1
#define IPX_CUBIC 2

/* periodic_mask helpers */
#define IPX_PERIODIC(axis) (1 << (axis))
#define IPX_WRAP_ONLY(axis) (1 << (4 + (axis)))

/* table layouts accepted by the evaluation kernels */
#define IPX_SOA 0 /* the reference's layout: 2^d separate C-order arrays (f, fx, ...)  */
#define IPX_AOS 1 /* one record per node written by ipx_pack*: see DESIGN.md "Data layout" */

/* slope stencils of approx_df (interpax/_fd_derivs.py:44-75) */
#define IPX_SLOPE_CUBIC 0       /* _cubic1      _fd_derivs.py:79-101  */
#define IPX_SLOPE_CUBIC2 1      /* _cubic2      _fd_derivs.py:160-300 */
#define IPX_SLOPE_CARDINAL 2    /* _cardinal    _fd_derivs.py:303-324 (catmull-rom: c = 0) */
#define IPX_SLOPE_MONOTONIC 3   /* _monotonic   _fd_derivs.py:327-387 */
#define IPX_SLOPE_MONOTONIC0 4  /* _monotonic, zero end slopes */
#define IPX_SLOPE_AKIMA 5       /* _akima       _fd_derivs.py:390-424 */
#define IPX_SLOPE_ZERO 6        /* nearest / linear -> zeros, _fd_derivs.py:72-73 */
#define IPX_SLOPE_AKIMA_COMPLEX 7 /* _akima on complex data: `inner` counts real elements, (re, im) pairs;
                                     the weights are moduli of complex differences (jnp.abs, :414) */

/* cubic2 boundary conditions (_fd_derivs.py:104-157) */
#define IPX_BC_NOT_A_KNOT 0
#define IPX_BC_FIRST 1  /* (1, value); "clamped" = value NULL (zeros) */
#define IPX_BC_SECOND 2 /* (2, value); "natural" = value NULL (zeros) */

/* Per-axis run-time operands (traced in the reference). */
typedef struct ipx_axis_params {
  int32_t derivative;  /* order along this axis; >= 4 gives zeros, < 0 acts as 0
                          (lax.switch clamps, _spline.py:1151)                         */
  int32_t keep_lo;     /* 1: extrap=True below x[0] (polynomial continues); 0: fill    */
  int32_t keep_hi;     /* same above x[-1]                                             */
  int32_t reserved;
  double fill_lo;      /* value written where q < x[0] when keep_lo == 0 (NaN for
                          extrap=False, the number for a numeric extrap)  _spline.py:1190-1220 */
  double fill_hi;
  double period;       /* |period| is used, and only on axes named in periodic_mask    */
} ipx_axis_params;

typedef struct ipx_params {
  ipx_axis_params axis[3];
} ipx_params;

typedef struct ipx_slope_opts {
  double c;                  /* cardinal tension, kwarg `c` (_fd_derivs.py:51)          */
  int32_t bc_start, bc_end;  /* IPX_BC_* for cubic2, kwarg `bc_type` (_fd_derivs.py:47) */
  const void* bc_start_value; /* device, outer*inner values of T, or NULL for zeros     */
  const void* bc_end_value;
} ipx_slope_opts;

int ipx_version(void);
/* Static description of the build: "sm_100a" etc.  Never NULL. */
const char* ipx_build_info(void);
/* Diagnostics: kernels enqueued so far by this process through the evaluation entry points
 * (ipx_eval*, ipx_eval_stencil_*, ipx_halo_table_*), skipped candidates of the device-side
 * choice included.  The one counter in the library; it influences nothing. */
uint64_t ipx_launch_count(void);

/* ---------------------------------------------------------------------------------
 * Query evaluation.  Replaces the bodies of
 *   interp1d  interpax/_spline.py:529-592   (search, gather, A_CUBIC contraction, extrap)
 *   interp2d  interpax/_spline.py:702-805
 *   interp3d  interpax/_spline.py:940-1105
 * including _get_t_der (:1138-1151), _extrap (:1180-1222) and, through the knot/perm
 * arguments, the table side of _make_periodic (:1108-1135) without copying any table.
 *
 *   q*          nq query coordinates per axis (already broadcast to a common length)
 *   knots*      n knots of that axis; on an axis named in periodic_mask, the n+2 padded,
 *               sorted knots written by ipx_periodic_knots_*
 *   perm*       NULL, or (periodic axes) n int32: sorted position -> table index
 *   n*          extent of the tables along the axis (un-padded)
 *   tables      HOST array of device pointers.  IPX_SOA: 2^d tables in the reference's
 *               stacking order (1-D: f, fx; 2-D: f, fx, fy, fxy; 3-D: f, fx, fy, fz, fxy,
 *               fxz, fyz, fxyz -- _spline.py:581-584, 779-783, 1070-1078), each of shape
 *               (nx[,ny[,nz]], batch) C-order; for IPX_NEAREST / IPX_LINEAR only
 *               tables[0] is read.  IPX_AOS: tables[0] is the packed array.
 *   batch       product of f's trailing dimensions (>= 1); out is (nq, batch)
 *   periodic_mask  bit a set: axis a is periodic (queries are wrapped with the Python
 *               remainder in the kernel; extrapolation flags of that axis are ignored,
 *               _spline.py:527).  Bit 4+a set instead (IPX_WRAP_ONLY): the caller has
 *               already materialised the padded knots AND padded tables of that axis
 *               (ipx_pad_periodic_*), n is the padded extent, and the kernel only wraps
 *               the query.
 *   params      see ipx_params; params_on_device != 0 means a device pointer
 *   out         (nq, batch) results
 *   idx_out     NULL or (d, nq) int32: the interval index i = clip(searchsorted(knots, q,
 *               side="right"), 1, len-1) per axis, in padded coordinates on periodic
 *               axes (_spline.py:571, 767-768, 1051-1053) -- for bit-exact index parity
 * --------------------------------------------------------------------------------- */
#define IPX_DECL_EVAL(SUF, T)                                                                   \
  int ipx_eval1d_##SUF(const T* xq, int64_t nq, const T* x, const int32_t* permx, int32_t nx,   \
                       const T* const* tables, int32_t layout, int64_t batch,                   \
                       int32_t method_class, int32_t periodic_mask, const ipx_params* params,   \
                       int32_t params_on_device, T* out, int32_t* idx_out, void* stream);       \
  int ipx_eval2d_##SUF(const T* xq, const T* yq, int64_t nq, const T* x, const T* y,            \
                       const int32_t* permx, const int32_t* permy, int32_t nx, int32_t ny,      \
                       const T* const* tables, int32_t layout, int64_t batch,                   \
                       int32_t method_class, int32_t periodic_mask, const ipx_params* params,   \
                       int32_t params_on_device, T* out, int32_t* idx_out, void* stream);       \
  int ipx_eval3d_##SUF(const T* xq, const T* yq, const T* zq, int64_t nq, const T* x,           \
                       const T* y, const T* z, const int32_t* permx, const int32_t* permy,      \
                       const int32_t* permz, int32_t nx, int32_t ny, int32_t nz,                \
                       const T* const* tables, int32_t layout, int64_t batch,                   \
                       int32_t method_class, int32_t periodic_mask, const ipx_params* params,   \
                       int32_t params_on_device, T* out, int32_t* idx_out, void* stream);

IPX_DECL_EVAL(f32, float)
IPX_DECL_EVAL(f64, double)

/* ---------------------------------------------------------------------------------
 * Cubic evaluation straight from f for the three-point slope stencils.
 * For method = cubic (_fd_derivs.py:79-101), cardinal and catmull-rom (:303-324) every link of
 * the slope chain fx, fy, fz, fxy, ... (interpax/_spline.py:380-393 classes, :1027-1040
 * functional form) is linear in f and acts along one axis, so
 *   interpNd(q, knots, f, method)      ==  sum over the 4^d neighbourhood of f, weights Tx (x) Ty (x) Tz
 * with four tap weights per axis (Hermite weights composed with the stencil).  These two entry
 * points replace the slope chain + gather + A_CUBIC/A_BICUBIC/A_TRICUBIC contraction
 * (_spline.py:572-589, 759-800, 1027-1099) and, through the mask bits, both periodic
 * conventions of the reference without materialising a padded table per call.
 *
 * ipx_halo_table_*: out = f (C order, extents n[ax], scalar-valued) copied into the halo layout,
 *   extent n[ax] + 2 on an open axis, n[ax] + 4 on a periodic one (mask bits as for
 *   ipx_eval_stencil_*; perm from ipx_periodic_knots_*, NULL = identity; perm itself a HOST array
 *   of ndim device pointers or NULL).  Entry u holds the value at lookup-knot position u - 1:
 *   open axis: f[u-1], and one GHOST row beyond each end, 2 f[0] - f[1] resp. 2 f[n-1] - f[n-2],
 *   with which the centred stencil at the end knot equals the reference's one-sided one;
 *   IPX_PERIODIC: f[perm[(u-2) mod n]] throughout; IPX_PAD_FIRST: the same for the padded
 *   positions 0 .. n+1 and ghost rows beyond them.
 * ipx_eval_stencil_*: q / knots / slope_knots are HOST arrays of ndim device pointers.
 *   periodic_mask bit a (IPX_PERIODIC): Interpolator* convention -- knots[a] are the n+2 padded
 *   sorted knots, slope_knots[a] the n original knots the slopes are taken on (one-sided at table
 *   rows 0 and n-1, then wrapped; the knots must already be sorted inside one period);
 *   bit 4+a (IPX_PAD_FIRST): functional-form convention -- slopes are taken on the padded knots
 *   (_spline.py:924-938 precede :1027-1040).  Open axes: knots[a] = the n knots.  slope_knots may
 *   be NULL when no axis uses IPX_PERIODIC.
 *   slope_method IPX_SLOPE_CUBIC or IPX_SLOPE_CARDINAL (c = tension; catmull-rom: c = 0).
 *   records: NULL, or the IPX_AOS records of the same tables (ipx_slopes_pack_* / ipx_pack_*; of the
 *   PADDED tables on IPX_PAD_FIRST axes).  With IPX_STENCIL_AUTO they serve the streams on which the
 *   record form measures faster (dense cell-sorted 3-D streams; random order on tables much larger
 *   than L2, and in 2-D).
 *   hint: IPX_STENCIL_SPARSE = one query per lane; IPX_STENCIL_DENSE = four consecutive queries per
 *   lane share the table rows, staged in shared memory (for cell-sorted streams with several
 *   queries per cell; any stream is still evaluated correctly); IPX_STENCIL_COLUMN (3-D only; other
 *   ranks run the sparse shape) = the rows of the CTA's cell columns staged in shared memory by bulk
 *   copies, queries re-paired by cell so a lane reads the 4^3 neighbourhood once for two queries (for
 *   dense cell-sorted streams; any stream is still evaluated correctly); IPX_STENCIL_AUTO chooses, and if
 *   `workspace` (NULL, or 4 int32 of device scratch) is given it chooses ON THE DEVICE: a probe
 *   kernel classifies a sample of the stream (dense cell-sorted / cell-sorted / random), every
 *   candidate kernel is enqueued and the ones not chosen return at once -- no host
 *   synchronisation, so the call stays CUDA-graph capturable.
 *   Returns IPX_EUNSUPPORTED when the knot vectors do not fit in shared memory
 *   (the caller then uses ipx_slopes_pack_* + ipx_eval*).
 *   params / out / idx_out / stream as for ipx_eval*.
 * --------------------------------------------------------------------------------- */
#define IPX_PAD_FIRST(axis) (1 << (4 + (axis)))
#define IPX_STENCIL_AUTO 0
#define IPX_STENCIL_SPARSE 1
#define IPX_STENCIL_DENSE 2
#define IPX_STENCIL_COLUMN 3
/* elements of the halo table (the last axis is padded to a multiple of 32 bytes), or -1 */
int64_t ipx_halo_table_elems(int32_t ndim, const int32_t* n, int32_t periodic_mask, int32_t elem_size);
int ipx_halo_table_f32(int32_t ndim, const float* f, const int32_t* n, int32_t periodic_mask,
                       const int32_t* const* perm, float* out, void* stream);
int ipx_halo_table_f64(int32_t ndim, const double* f, const int32_t* n, int32_t periodic_mask,
                       const int32_t* const* perm, double* out, void* stream);
int ipx_eval_stencil_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                         const float* const* slope_knots, const int32_t* n, const float* halo, const float* records,
                         int32_t slope_method, double c, int32_t periodic_mask, const ipx_params* params,
                         int32_t params_on_device, float* out, int32_t* idx_out, int32_t hint,
                         int32_t* workspace, void* stream);
int ipx_eval_stencil_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                         const double* const* slope_knots, const int32_t* n, const double* halo, const double* records,
                         int32_t slope_method, double c, int32_t periodic_mask, const ipx_params* params,
                         int32_t params_on_device, double* out, int32_t* idx_out, int32_t hint,
                         int32_t* workspace, void* stream);

/* ---------------------------------------------------------------------------------
 * Interval lookup alone: idx[k] = clip(searchsorted(knots, q[k], side="right"), 1, n-1).
 * Replaces the jnp.searchsorted + jnp.clip pair at interpax/_spline.py:543,556,571,
 * 708-709,729-730,767-768,946-948,983-985,1051-1053.
 * --------------------------------------------------------------------------------- */
int ipx_bin_index_f32(const float* knots, int32_t n, const float* q, int64_t nq, int32_t* idx,
                      void* stream);
int ipx_bin_index_f64(const double* knots, int32_t n, const double* q, int64_t nq, int32_t* idx,
                      void* stream);

/* ---------------------------------------------------------------------------------
 * Slope precompute: out = approx_df(x, f, method, axis) for a C-contiguous table viewed
 * as (outer, n, inner) with the differentiated axis in the middle.
 * Replaces interpax/_fd_derivs.py:9-443 (one call per link of the chains at
 * _spline.py:119-122, 237-244, 380-393, 572-573, 759-764, 1027-1040).
 * `workspace` must hold ipx_slopes_workspace_bytes(...) bytes (device memory).
 * --------------------------------------------------------------------------------- */
size_t ipx_slopes_workspace_bytes(int32_t method, int32_t n, int32_t elem_size);
int ipx_slopes_f32(const float* x, int32_t n, const float* f, int64_t outer, int64_t inner,
                   int32_t method, const ipx_slope_opts* opts, float* out, void* workspace,
                   void* stream);
int ipx_slopes_f64(const double* x, int32_t n, const double* f, int64_t outer, int64_t inner,
                   int32_t method, const ipx_slope_opts* opts, double* out, void* workspace,
                   void* stream);

/* ---------------------------------------------------------------------------------
 * Reverse mode of the slope precompute with respect to f, for the methods that are linear in f:
 * grad_f (+)= (d approx_df(x, f, method, axis) / d f)^T g, same (outer, n, inner) view.  This is the
 * rule jax.grad derives by tracing interpax/_fd_derivs.py:79-101 (cubic), :303-324 (cardinal) and
 * :160-300 (cubic2: a transposed tridiagonal solve; boundary VALUES do not depend on f), and what a
 * gradient with respect to f of any cubic2 / cubic / cardinal interpolant chains behind
 * ipx_vjp_tables_* (one call per link of the chain, in reverse).  accumulate != 0 adds into grad_f.
 * g and grad_f must not alias.  IPX_SLOPE_MONOTONIC* / AKIMA* (not linear in f) and the two- /
 * three-knot special systems of cubic2: IPX_EUNSUPPORTED.
 * `workspace`: ipx_slopes_vjp_workspace_bytes(...) bytes of device memory.
 * --------------------------------------------------------------------------------- */
size_t ipx_slopes_vjp_workspace_bytes(int32_t method, int32_t n, int64_t outer, int64_t inner,
                                      int32_t elem_size);
int ipx_slopes_vjp_f32(const float* x, int32_t n, const float* g, int64_t outer, int64_t inner,
                       int32_t method, const ipx_slope_opts* opts, int32_t accumulate, float* grad_f,
                       void* workspace, void* stream);
int ipx_slopes_vjp_f64(const double* x, int32_t n, const double* g, int64_t outer, int64_t inner,
                       int32_t method, const ipx_slope_opts* opts, int32_t accumulate, double* grad_f,
                       void* workspace, void* stream);

/* ---------------------------------------------------------------------------------
 * Periodic axis preparation -- the knot side of _make_periodic, interpax/_spline.py:
 * 1117-1122: xm = x % |period|; perm = stable argsort(xm); xpad = [xs[-1]-P, xs..., xs[0]+P].
 * xpad has n+2 entries, perm n.  `workspace` holds (n + 4) * elem_size bytes.
 * --------------------------------------------------------------------------------- */
int ipx_periodic_knots_f32(const float* x, int32_t n, double period, float* xpad, int32_t* perm,
                           void* workspace, void* stream);
int ipx_periodic_knots_f64(const double* x, int32_t n, double period, double* xpad,
                           int32_t* perm, void* workspace, void* stream);

/* Materialised wrap-pad of one table along one axis, for the functional form that
 * differentiates *after* padding (interpax/_spline.py:1123-1135 followed by :1027-1040):
 * out[o, p, i] = in[o, perm[(p-1) mod n], i], p in [0, n+2). */
int ipx_pad_periodic_f32(const float* in, int64_t outer, int32_t n, int64_t inner,
                         const int32_t* perm, float* out, void* stream);
int ipx_pad_periodic_f64(const double* in, int64_t outer, int32_t n, int64_t inner,
                         const int32_t* perm, double* out, void* stream);

/* ---------------------------------------------------------------------------------
 * Interleave the 2^d SoA tables into the IPX_AOS layout (one record of 2^d values per
 * (node, batch element)); record order in DESIGN.md.  `tables` is a HOST array of 2^d
 * device pointers in the reference's stacking order; nodes = nx*ny*nz.
 * --------------------------------------------------------------------------------- */
int ipx_pack_f32(int32_t ndim, const float* const* tables, int64_t nodes_times_batch, float* out,
                 void* stream);
int ipx_pack_f64(int32_t ndim, const double* const* tables, int64_t nodes_times_batch,
                 double* out, void* stream);

/* ---------------------------------------------------------------------------------
 * Fused knot slopes + pack: the whole chain of Interpolator2D/3D.__init__
 * (interpax/_spline.py:232-246, 380-393 -- fx, fy[, fz] from f; fxy from fx; fxz, fyz;
 * fxyz from fxy) in ONE pass from f (C order, shape (n[0], .., n[ndim-1], batch)) to the
 * IPX_AOS records, bit-identical to the per-table slopes followed by the pack above, but
 * without materialising the 2^d - 1 intermediate tables.  knots / n are HOST arrays of ndim
 * entries (device pointers / extents).  Three-point stencils only: IPX_SLOPE_CUBIC, _CARDINAL (opts->c), _MONOTONIC,
 * _MONOTONIC0, ndim 2 or 3 and every n >= 3; anything else returns IPX_EUNSUPPORTED and the
 * caller uses ipx_slopes_* + ipx_pack_*.
 * --------------------------------------------------------------------------------- */
int ipx_slopes_pack_f32(int32_t ndim, const float* const* knots, const int32_t* n, const float* f,
                        int64_t batch, int32_t method, const ipx_slope_opts* opts, float* out,
                        void* stream);
int ipx_slopes_pack_f64(int32_t ndim, const double* const* knots, const int32_t* n,
                        const double* f, int64_t batch, int32_t method, const ipx_slope_opts* opts,
                        double* out, void* stream);

/* ---------------------------------------------------------------------------------
 * Spatial pre-sort of the query stream (no counterpart in the reference, which evaluates
 * queries in the order given, interpax/_spline.py:1051-1099).
 * ipx_cell_order_*: order[k] = index of the k-th query in cell order (stable argsort of the
 * linear cell key built from the same interval lookup as the evaluation).  q / knots are
 * HOST arrays of ndim device pointers, n the table extents, periodic_mask / params as for
 * ipx_eval*.  nq < 2^32, number of cells < 2^32.  workspace: ipx_cell_order_workspace_bytes(nq).
 * ipx_gather_*:  dst[k*batch + b] = src[order[k]*batch + b];  ipx_scatter_*: the inverse.
 * A caller evaluates  scatter(eval(gather(q)))  and gets bit-identical results to eval(q).
 * --------------------------------------------------------------------------------- */
size_t ipx_cell_order_workspace_bytes(int64_t nq);
int ipx_cell_order_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                       const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                       int32_t params_on_device, uint32_t* order, void* workspace, void* stream);
int ipx_cell_order_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                       const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                       int32_t params_on_device, uint32_t* order, void* workspace, void* stream);
/* ipx_cell_sort_*: ipx_cell_order_* plus the queries themselves in that order, q_sorted[ax][k] =
 * q[ax][order[k]] (HOST array of ndim device pointers, nq elements each).  The coordinates
 * ride through an interleaved copy in the workspace so the permutation reads one sector per
 * query.  workspace: ipx_cell_sort_workspace_bytes(ndim, nq, sizeof(T)). */
size_t ipx_cell_sort_workspace_bytes(int32_t ndim, int64_t nq, int32_t elem_size);
int ipx_cell_sort_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                      const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                      int32_t params_on_device, uint32_t* order, float* const* q_sorted,
                      void* workspace, void* stream);
int ipx_cell_sort_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                      const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                      int32_t params_on_device, uint32_t* order, double* const* q_sorted,
                      void* workspace, void* stream);
/* ipx_cell_order_records_*: ipx_cell_order_* plus the interleaved copy of the coordinates it sorts by,
 * records[e * R + ax] = q[ax][e] with R = 1, 2, 4 for ndim = 1, 2, 3 (32-byte aligned, nq * R
 * elements, written for the caller's use): what ipx_eval_stencil_perm_* reads, one sector per query.
 * workspace: ipx_cell_order_workspace_bytes(nq). */
int ipx_cell_order_records_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                               const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                               int32_t params_on_device, uint32_t* order, float* records, void* workspace,
                               void* stream);
int ipx_cell_order_records_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                               const int32_t* n, int32_t periodic_mask, const ipx_params* params,
                               int32_t params_on_device, uint32_t* order, double* records, void* workspace,
                               void* stream);
/* ipx_eval_stencil_perm_*: ipx_eval_stencil_* visiting the queries in the order given --
 * out[order[k]] = value at query order[k], k = 0 .. nq-1, coordinates read from coord_records -- i.e.
 * scatter(eval(gather(q))) in ONE kernel: the pre-sorted call without its two permutation passes.
 * Results are bit-identical to ipx_eval_stencil_* on the same queries.  q is still passed (argument
 * checks only).  One query per lane (the sparse shape). */
int ipx_eval_stencil_perm_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                              const float* const* slope_knots, const int32_t* n, const float* halo,
                              int32_t slope_method, double c, int32_t periodic_mask, const ipx_params* params,
                              int32_t params_on_device, const uint32_t* order, const float* coord_records,
                              float* out, void* stream);
int ipx_eval_stencil_perm_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                              const double* const* slope_knots, const int32_t* n, const double* halo,
                              int32_t slope_method, double c, int32_t periodic_mask, const ipx_params* params,
                              int32_t params_on_device, const uint32_t* order, const double* coord_records,
                              double* out, void* stream);
int ipx_gather_f32(const float* src, const uint32_t* order, int64_t n, int64_t batch, float* dst,
                   void* stream);
int ipx_gather_f64(const double* src, const uint32_t* order, int64_t n, int64_t batch, double* dst,
                   void* stream);
/* up to 3 arrays at once (the query coordinates): dst[a][k] = src[a][order[k]] */
int ipx_gather_multi_f32(int32_t narrays, const float* const* src, const uint32_t* order, int64_t n,
                         float* const* dst, void* stream);
int ipx_gather_multi_f64(int32_t narrays, const double* const* src, const uint32_t* order, int64_t n,
                         double* const* dst, void* stream);
int ipx_scatter_f32(const float* src, const uint32_t* order, int64_t n, int64_t batch, float* dst,
                    void* stream);
int ipx_scatter_f64(const double* src, const uint32_t* order, int64_t n, int64_t batch, double* dst,
                    void* stream);

/* ---------------------------------------------------------------------------------
 * Reverse-mode rule with respect to the tables: grad_tables[ff][node, b] += cotangent[q, b] *
 * d out[q, b] / d tables[ff][node, b] (accumulates: zero the gradients first).  This is the
 * transpose of the gather + A @ F contraction that JAX derives automatically in the reference
 * (interpax/_spline.py:581-589, 779-800, 1070-1099).  q / knots / perm are HOST arrays of ndim
 * device pointers; grad_tables a HOST array of 2^ndim device pointers in the reference's
 * stacking order (IPX_CUBIC) or of one pointer (IPX_LINEAR).  IPX_NEAREST: IPX_EUNSUPPORTED.
 * Queries replaced by an extrapolation fill contribute nothing.
 * --------------------------------------------------------------------------------- */
int ipx_vjp_tables_f32(int32_t ndim, const float* const* q, int64_t nq, const float* const* knots,
                       const int32_t* const* perm, const int32_t* n, int64_t batch,
                       int32_t method_class, int32_t periodic_mask, const ipx_params* params,
                       int32_t params_on_device, const float* cotangent, float* const* grad_tables,
                       void* stream);
int ipx_vjp_tables_f64(int32_t ndim, const double* const* q, int64_t nq, const double* const* knots,
                       const int32_t* const* perm, const int32_t* n, int64_t batch,
                       int32_t method_class, int32_t periodic_mask, const ipx_params* params,
                       int32_t params_on_device, const double* cotangent, double* const* grad_tables,
                       void* stream);

/* ---------------------------------------------------------------------------------
 * Derivatives of the cubic evaluation with respect to the KNOTS and to f, for the three-point
 * stencils (IPX_SLOPE_CUBIC, IPX_SLOPE_CARDINAL): the map (knots, f) -> interpNd(q, knots, f) with
 * the slope chain inside, which is what the reference's AD tests differentiate
 * (tests/test_interpolate.py:542-637: jacfwd / jacrev with respect to the knot vector; JAX derives
 * it by tracing interpax/_spline.py:1027-1099 and _fd_derivs.py:79-101, 303-324).
 * Arguments as for ipx_eval_stencil_* except that `f` is the ORIGINAL table (C order, extents n,
 * scalar-valued), perm is a HOST array of ndim device pointers or NULL (IPX_PERIODIC axes need
 * knots sorted inside one period, perm NULL), and:
 *   ipx_vjp_stencil_*  reverse mode: grad_knots[ax][k] += sum_q cotangent[q] d out[q]/d x_ax[k]
 *                      (HOST array of ndim device pointers, n[ax] entries each, entries may be NULL),
 *                      grad_f += the same with respect to f (or NULL).  ACCUMULATES: zero first.
 *   ipx_jvp_stencil_*  forward mode: tangent[q] = sum_k knots_dot[ax][k] d out[q]/d x_ax[k]
 *                      + sum f_dot * d out[q]/d f   (either may be NULL = zero tangent).
 * Gradients are with respect to the caller's knot vectors (on periodic axes: the n original
 * knots).  Queries replaced by an extrapolation fill have zero derivative.  Other slope
 * methods: IPX_EUNSUPPORTED.
 * --------------------------------------------------------------------------------- */
#define IPX_DECL_GRAD(SUF, T)                                                                             \
  int ipx_vjp_stencil_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,           \
                            const T* const* slope_knots, const int32_t* const* perm, const int32_t* n,    \
                            const T* f, int32_t slope_method, double c, int32_t periodic_mask,            \
                            const ipx_params* params, int32_t params_on_device, const T* cotangent,       \
                            T* const* grad_knots, T* grad_f, void* stream);                               \
  int ipx_jvp_stencil_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,           \
                            const T* const* slope_knots, const int32_t* const* perm, const int32_t* n,    \
                            const T* f, int32_t slope_method, double c, int32_t periodic_mask,            \
                            const ipx_params* params, int32_t params_on_device,                           \
                            const T* const* knots_dot, const T* f_dot, T* tangent, void* stream);
IPX_DECL_GRAD(f32, float)
IPX_DECL_GRAD(f64, double)

/* ---------------------------------------------------------------------------------
 * Reverse mode with respect to the knots for the TABLE form of the cubic evaluation and for the
 * slope precompute -- the pieces from which the knot gradient of cubic2 (a global solve) and
 * monotonic (not linear in f) interpolants is chained, the two methods of the reference's TestAD
 * (tests/test_interpolate.py:542-637) that ipx_vjp_stencil_* does not cover.
 *
 * ipx_vjp_knots_tables_*: grad_knots[ax][k] += sum_q cotangent[q] * d out[q] / d knots[ax][k] with the
 *   2^ndim tables held FIXED (the dependence of _spline.py:576-589, 770-800, 1055-1099 on the cell's
 *   two knots: local coordinate and the dx that scales the slope tables).  Arguments as for
 *   ipx_vjp_tables_* (IPX_CUBIC, scalar-valued tables in the reference's stacking order, SoA);
 *   grad_knots: HOST array of ndim device pointers (entries may be NULL), each over the nk LOOKUP
 *   knots of its axis (n, or the n + 2 padded knots of a periodic axis).  Accumulates.
 * ipx_slopes_knots_vjp_*: one link of the slope chain, slopes = approx_df(x, f, method) along the
 *   middle axis of the (outer, n, inner) view: grad_x[k] += sum y * d slopes / d x[k]; for
 *   IPX_SLOPE_MONOTONIC* also grad_f += (d slopes / d f)^T y (that method is not linear in f, so
 *   ipx_slopes_vjp_* does not serve it).  y is the cotangent of the slopes, except for
 *   IPX_SLOPE_CUBIC2, where it is T^-T applied to that cotangent -- the array ipx_slopes_vjp_*
 *   leaves in its workspace at element offset 3 n -- and `slopes` must hold the forward result of
 *   the link (NULL otherwise).  opts as in the forward call (boundary values included).  n >= 3.
 * --------------------------------------------------------------------------------- */
/* Forward mode of the same pieces (jax.jacfwd in the reference's TestAD):
 * ipx_jvp_knots_tables_*: tangent[q] = sum_k knots_dot[ax][k] * d out[q] / d knots[ax][k], tables fixed
 *   (knots_dot over the lookup knots, entries may be NULL); the tangent of the tables is added by an
 *   ordinary ipx_eval*d_* call on the tangent tables (the evaluation is linear in them).
 * ipx_slopes_jvp_rows_*: directional derivative of every row of one link along (x_dot, f_dot) (either may
 *   be NULL): the tangent of the slopes for cubic / cardinal / monotonic; for IPX_SLOPE_CUBIC2 the
 *   right-hand side b' - T' s of the tangent system, which ipx_cubic2_solve_* then solves in place:
 * ipx_cubic2_solve_*: y <- T(x)^-1 y for the spline matrix of _fd_derivs.py:199-286 (boundary KINDS
 *   IPX_BC_*), Thomas without pivoting along the middle axis of the (outer, n, inner) view; workspace:
 *   ipx_slopes_workspace_bytes(IPX_SLOPE_CUBIC2, n, sizeof(T)). */
#define IPX_DECL_GRAD_TABLES(SUF, T)                                                                      \
  int ipx_jvp_knots_tables_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,      \
                                 const int32_t* const* perm, const int32_t* n, const T* const* tables,    \
                                 int32_t periodic_mask, const ipx_params* params,                         \
                                 int32_t params_on_device, const T* const* knots_dot, T* tangent,         \
                                 void* stream);                                                           \
  int ipx_slopes_jvp_rows_##SUF(const T* x, int32_t n, const T* f, const T* slopes, const T* x_dot,       \
                                const T* f_dot, int64_t outer, int64_t inner, int32_t method,             \
                                const ipx_slope_opts* opts, T* out, void* stream);                        \
  int ipx_cubic2_solve_##SUF(const T* x, int32_t n, int32_t bc_start, int32_t bc_end, T* y,               \
                             int64_t outer, int64_t inner, void* workspace, void* stream);                \
  int ipx_vjp_knots_tables_##SUF(int32_t ndim, const T* const* q, int64_t nq, const T* const* knots,      \
                                 const int32_t* const* perm, const int32_t* n, const T* const* tables,    \
                                 int32_t periodic_mask, const ipx_params* params,                         \
                                 int32_t params_on_device, const T* cotangent, T* const* grad_knots,      \
                                 void* stream);                                                           \
  int ipx_slopes_knots_vjp_##SUF(const T* x, int32_t n, const T* f, const T* slopes, const T* y,          \
                                 int64_t outer, int64_t inner, int32_t method, const ipx_slope_opts* opts,\
                                 T* grad_x, T* grad_f, void* stream);
IPX_DECL_GRAD_TABLES(f32, float)
IPX_DECL_GRAD_TABLES(f64, double)

/* ---------------------------------------------------------------------------------
 * Piecewise polynomials in the power basis (interpax/_ppoly.py).  Coefficients as the
 * reference holds them after its moveaxis: c[(j*m + i)*batch + b] multiplies
 * (x - x[i])^(k-1-j) on piece i, for trailing element b; x has m+1 breakpoints.
 *
 * ipx_ppoly_eval_*: PPoly.__call__ (_ppoly.py:189-257): out[q*batch + b] = nu-th derivative of
 *   piece clip(searchsorted(x, q, "right"), 1, m) at q (zero when nu >= k).
 *   layout: IPX_SOA = the (k, m, batch) array above; IPX_AOS = the piece-major copy (m, k)
 *   written by ipx_ppoly_pack_* (batch == 1 only: the k coefficients of a piece share a sector,
 *   which is what random-order queries pay for).
 *   extrap: IPX_PPOLY_NAN (NaN outside [x[0], x[m]]), IPX_PPOLY_EXTRAPOLATE (end pieces
 *   extended), IPX_PPOLY_PERIODIC (q -> x[0] + (q - x[0]) mod (x[m] - x[0]) first).
 *   idx_out: optional, the piece index + 1 (int32, nq).  k <= 64.
 * ipx_ppoly_from_hermite_*: CubicHermiteSpline.__init__ (_ppoly.py:540-569): values y and
 *   slopes dydx at the n knots, shape (n, batch), -> c of shape (4, n-1, batch).
 * ipx_ppoly_derivative_*: coefficients of PPoly.derivative(nu) (_ppoly.py:259-296), 0 <= nu < k:
 *   c2 of shape (k-nu, m, batch).
 * ipx_ppoly_antiderivative_*: ONE step of PPoly.antiderivative (_ppoly.py:328-339): c2 of shape
 *   (k+1, m, batch), continuous across the breakpoints (constants accumulated left to right).
 * --------------------------------------------------------------------------------- */
#define IPX_PPOLY_NAN 0
#define IPX_PPOLY_EXTRAPOLATE 1
#define IPX_PPOLY_PERIODIC 2
int ipx_ppoly_eval_f32(const float* c, int32_t layout, int32_t k, int32_t m, int64_t batch,
                       const float* x, const float* q, int64_t nq, int32_t nu, int32_t extrap,
                       float* out, int32_t* idx_out, void* stream);
int ipx_ppoly_eval_f64(const double* c, int32_t layout, int32_t k, int32_t m, int64_t batch,
                       const double* x, const double* q, int64_t nq, int32_t nu, int32_t extrap,
                       double* out, int32_t* idx_out, void* stream);
int ipx_ppoly_pack_f32(const float* c, int32_t k, int32_t m, float* out, void* stream);
int ipx_ppoly_pack_f64(const double* c, int32_t k, int32_t m, double* out, void* stream);
int ipx_ppoly_from_hermite_f32(const float* x, int32_t n, const float* y, const float* dydx,
                               int64_t batch, float* c, void* stream);
int ipx_ppoly_from_hermite_f64(const double* x, int32_t n, const double* y, const double* dydx,
                               int64_t batch, double* c, void* stream);
int ipx_ppoly_derivative_f32(const float* c, int32_t k, int32_t m, int64_t batch, int32_t nu,
                             float* c2, void* stream);
int ipx_ppoly_derivative_f64(const double* c, int32_t k, int32_t m, int64_t batch, int32_t nu,
                             double* c2, void* stream);
int ipx_ppoly_antiderivative_f32(const float* c, int32_t k, int32_t m, int64_t batch,
                                 const float* x, float* c2, void* stream);
int ipx_ppoly_antiderivative_f64(const double* c, int32_t k, int32_t m, int64_t batch,
                                 const double* x, double* c2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IPX_H_ */
