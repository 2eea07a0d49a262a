// xla_ffi_shim.cc -- XLA FFI handlers that forward to the C ABI of libipx.so: one handler per
// family of entry points of include/ipx.h (float32 / float64 dispatched on the operand type,
// 1-D / 2-D / 3-D on the `ndim` attribute), so that interpax_b200/jax_frontend.py can express
// interp1d/2d/3d, Interpolator1D/2D/3D and approx_df as XLA custom calls.
//
// NOT BUILT IN THIS ENVIRONMENT: jaxlib (which ships xla/ffi/api/ffi.h, see
// jax.ffi.include_dir()) is not installed here and cannot be, so this file has never been
// compiled or run.  It is written against the documented XLA FFI API (jax >= 0.4.38) and
// compiles to nothing when the header is absent.
//
// Build (where JAX is available):
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include interpax_b200/csrc/xla_ffi_shim.cc \
//       -L interpax_b200 -lipx -lcudart -o interpax_b200/libipx_xla.so
//
// Conventions shared by the handlers:
//   * every operand / result is a device buffer owned by XLA; the stream is XLA's; nothing is
//     allocated here -- scratch memory arrives as extra RESULT buffers that the Python side drops;
//   * run-time operands of the reference (derivative orders, extrap flags and fill values,
//     periods: only `method` is static under jit, interpax/_spline.py:444,595,808) travel as the
//     120 bytes of an `ipx_params` in a uint8 device buffer;
//   * batching (jax.vmap with vmap_method="expand_dims"): queries may carry leading batch
//     dimensions, which simply extend the query count; tables either carry none (size-1 leading
//     dimensions) or the same ones, in which case the handler loops over them.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define IPX_HAVE_XLA_FFI 1
#endif
#endif

#ifdef IPX_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>
#include <vector>

#include "ipx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error to_error(int rc, const char* what) {
  if (rc == IPX_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + " failed with code " + std::to_string(rc));
}
ffi::Error bad(const std::string& msg) { return ffi::Error(ffi::ErrorCode::kInvalidArgument, msg); }

struct Args {
  ffi::RemainingArgs& a;
  size_t next = 0;
  ffi::Error err = ffi::Error::Success();
  bool ok = true;
  ffi::AnyBuffer take() {
    auto b = a.get<ffi::AnyBuffer>(next++);
    if (!b.has_value()) {
      ok = false;
      err = b.error();
      return a.get<ffi::AnyBuffer>(0).value();
    }
    return b.value();
  }
};

int64_t product(ffi::Span<const int64_t> d, size_t lo, size_t hi) {
  int64_t p = 1;
  for (size_t i = lo; i < hi; ++i) p *= d[i];
  return p;
}
const void* data_or_null(const ffi::AnyBuffer& b) { return b.element_count() == 0 ? nullptr : b.untyped_data(); }

// ---------------------------------------------------------------------------------------------
// ipx_eval{1,2,3}d_{f32,f64}: operands q[ndim], knots[ndim], perm[ndim] (int32, empty = none),
// params (uint8[120]), tables[ntab]; result out (..., nq, batch).
// ---------------------------------------------------------------------------------------------
template <typename T>
int call_eval(int ndim, const T* const* q, int64_t nq, const T* const* k, const int32_t* const* p, const int32_t* n,
              const T* const* tabs, int32_t layout, int64_t batch, int32_t mclass, int32_t mask,
              const ipx_params* params, T* out, void* s);
template <>
int call_eval<float>(int ndim, const float* const* q, int64_t nq, const float* const* k, const int32_t* const* p,
                     const int32_t* n, const float* const* tabs, int32_t layout, int64_t batch, int32_t mclass,
                     int32_t mask, const ipx_params* params, float* out, void* s) {
  if (ndim == 1) return ipx_eval1d_f32(q[0], nq, k[0], p[0], n[0], tabs, layout, batch, mclass, mask, params, 1, out, nullptr, s);
  if (ndim == 2)
    return ipx_eval2d_f32(q[0], q[1], nq, k[0], k[1], p[0], p[1], n[0], n[1], tabs, layout, batch, mclass, mask, params,
                          1, out, nullptr, s);
  return ipx_eval3d_f32(q[0], q[1], q[2], nq, k[0], k[1], k[2], p[0], p[1], p[2], n[0], n[1], n[2], tabs, layout, batch,
                        mclass, mask, params, 1, out, nullptr, s);
}
template <>
int call_eval<double>(int ndim, const double* const* q, int64_t nq, const double* const* k, const int32_t* const* p,
                      const int32_t* n, const double* const* tabs, int32_t layout, int64_t batch, int32_t mclass,
                      int32_t mask, const ipx_params* params, double* out, void* s) {
  if (ndim == 1) return ipx_eval1d_f64(q[0], nq, k[0], p[0], n[0], tabs, layout, batch, mclass, mask, params, 1, out, nullptr, s);
  if (ndim == 2)
    return ipx_eval2d_f64(q[0], q[1], nq, k[0], k[1], p[0], p[1], n[0], n[1], tabs, layout, batch, mclass, mask, params,
                          1, out, nullptr, s);
  return ipx_eval3d_f64(q[0], q[1], q[2], nq, k[0], k[1], k[2], p[0], p[1], p[2], n[0], n[1], n[2], tabs, layout, batch,
                        mclass, mask, params, 1, out, nullptr, s);
}

template <typename T>
ffi::Error eval_typed(cudaStream_t stream, ffi::RemainingArgs& ra, ffi::AnyBuffer& out, int32_t ndim, int32_t layout,
                      int32_t mclass, int32_t mask, int64_t batch, int32_t ntab) {
  Args A{ra};
  std::vector<ffi::AnyBuffer> q, k, p, t;
  for (int i = 0; i < ndim; ++i) q.push_back(A.take());
  for (int i = 0; i < ndim; ++i) k.push_back(A.take());
  for (int i = 0; i < ndim; ++i) p.push_back(A.take());
  ffi::AnyBuffer params = A.take();
  for (int i = 0; i < ntab; ++i) t.push_back(A.take());
  if (!A.ok) return A.err;
  if (params.size_bytes() < sizeof(ipx_params)) return bad("params must hold an ipx_params");
  // table extents = the `ndim` dimensions in front of the trailing batch dimensions (AoS: and of
  // the record dimension); anything in front of those is a vmap batch
  const auto td = t[0].dimensions();
  int64_t tail = 1, used = 0;
  size_t pos = td.size();
  while (pos > 0 && tail < batch * (layout == IPX_AOS ? (1 << ndim) : 1)) tail *= td[--pos];
  if (pos < (size_t)ndim) return bad("table rank too small");
  int32_t n[3] = {0, 0, 0};
  for (int i = 0; i < ndim; ++i) n[i] = (int32_t)td[pos - ndim + i];
  const int64_t tbatch = product(td, 0, pos - ndim);  // vmap batch of the tables (1 = shared)
  const int64_t nq_all = (int64_t)q[0].element_count();
  if (tbatch != 1 && nq_all % tbatch != 0) return bad("batched tables need equally batched queries");
  const int64_t nq = nq_all / tbatch;
  const int64_t tstride = (int64_t)t[0].element_count() / tbatch;
  (void)used;
  for (int64_t b = 0; b < tbatch; ++b) {
    const T* qp[3];
    const T* kp[3];
    const int32_t* pp[3];
    const T* tp[8];
    for (int i = 0; i < ndim; ++i) {
      qp[i] = static_cast<const T*>(q[i].untyped_data()) + b * nq;
      kp[i] = static_cast<const T*>(k[i].untyped_data());
      pp[i] = static_cast<const int32_t*>(data_or_null(p[i]));
    }
    for (int i = 0; i < ntab; ++i) tp[i] = static_cast<const T*>(t[i].untyped_data()) + b * tstride;
    int rc = call_eval<T>(ndim, qp, nq, kp, pp, n, tp, layout, batch, mclass, mask,
                          static_cast<const ipx_params*>(params.untyped_data()),
                          static_cast<T*>(out.untyped_data()) + b * nq * batch, stream);
    if (rc != IPX_OK) return to_error(rc, "ipx_eval");
  }
  return ffi::Error::Success();
}

ffi::Error EvalImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::Result<ffi::AnyBuffer> out, int32_t ndim,
                    int32_t layout, int32_t method_class, int32_t periodic_mask, int64_t batch, int32_t ntab) {
  if (ndim < 1 || ndim > 3 || ntab < 1 || ntab > 8) return bad("ndim / ntab out of range");
  if (out->element_type() == ffi::F64)
    return eval_typed<double>(stream, args, *out, ndim, layout, method_class, periodic_mask, batch, ntab);
  if (out->element_type() == ffi::F32)
    return eval_typed<float>(stream, args, *out, ndim, layout, method_class, periodic_mask, batch, ntab);
  return bad("ipx_eval: float32 or float64 only");
}

// ---------------------------------------------------------------------------------------------
// ipx_eval_stencil_*: operands q[ndim], knots[ndim], slope_knots[ndim] (empty = none), params,
// halo, records (empty = none); results out (..., nq, 1), workspace int32[4].
// ---------------------------------------------------------------------------------------------
template <typename T>
ffi::Error stencil_typed(cudaStream_t stream, ffi::RemainingArgs& ra, ffi::AnyBuffer& out, ffi::AnyBuffer& ws,
                         int32_t ndim, int32_t slope_method, double c, int32_t mask, int32_t hint,
                         ffi::Span<const int64_t> n64) {
  Args A{ra};
  std::vector<ffi::AnyBuffer> q, k, sk;
  for (int i = 0; i < ndim; ++i) q.push_back(A.take());
  for (int i = 0; i < ndim; ++i) k.push_back(A.take());
  for (int i = 0; i < ndim; ++i) sk.push_back(A.take());
  ffi::AnyBuffer params = A.take(), halo = A.take(), rec = A.take();
  if (!A.ok) return A.err;
  const T* qp[3];
  const T* kp[3];
  const T* sp[3];
  int32_t n[3] = {0, 0, 0};
  bool any_sk = false;
  for (int i = 0; i < ndim; ++i) {
    qp[i] = static_cast<const T*>(q[i].untyped_data());
    kp[i] = static_cast<const T*>(k[i].untyped_data());
    sp[i] = static_cast<const T*>(data_or_null(sk[i]));
    any_sk = any_sk || sp[i] != nullptr;
    n[i] = (int32_t)n64[i];
  }
  const T* recp = static_cast<const T*>(data_or_null(rec));
  int rc;
  if constexpr (sizeof(T) == 8)
    rc = ipx_eval_stencil_f64(ndim, qp, (int64_t)q[0].element_count(), kp, any_sk ? sp : nullptr, n,
                              static_cast<const double*>(halo.untyped_data()), recp, slope_method, c, mask,
                              static_cast<const ipx_params*>(params.untyped_data()), 1,
                              static_cast<double*>(out.untyped_data()), nullptr, hint,
                              static_cast<int32_t*>(ws.untyped_data()), stream);
  else
    rc = ipx_eval_stencil_f32(ndim, qp, (int64_t)q[0].element_count(), kp, any_sk ? sp : nullptr, n,
                              static_cast<const float*>(halo.untyped_data()), recp, slope_method, c, mask,
                              static_cast<const ipx_params*>(params.untyped_data()), 1,
                              static_cast<float*>(out.untyped_data()), nullptr, hint,
                              static_cast<int32_t*>(ws.untyped_data()), stream);
  return to_error(rc, "ipx_eval_stencil");
}

ffi::Error StencilImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::Result<ffi::AnyBuffer> out,
                       ffi::Result<ffi::AnyBuffer> ws, int32_t ndim, int32_t slope_method, double c,
                       int32_t periodic_mask, int32_t hint, ffi::Span<const int64_t> n) {
  if (ndim < 1 || ndim > 3 || (int)n.size() != ndim) return bad("ndim / n mismatch");
  if (out->element_type() == ffi::F64)
    return stencil_typed<double>(stream, args, *out, *ws, ndim, slope_method, c, periodic_mask, hint, n);
  if (out->element_type() == ffi::F32)
    return stencil_typed<float>(stream, args, *out, *ws, ndim, slope_method, c, periodic_mask, hint, n);
  return bad("ipx_eval_stencil: float32 or float64 only");
}

// ---------------------------------------------------------------------------------------------
// ipx_halo_table_*: operands f, perm[ndim] (empty = identity); result halo (1-D, ipx_halo_table_elems)
// ---------------------------------------------------------------------------------------------
ffi::Error HaloImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::Result<ffi::AnyBuffer> out, int32_t ndim,
                    int32_t periodic_mask) {
  Args A{args};
  ffi::AnyBuffer f = A.take();
  std::vector<ffi::AnyBuffer> p;
  for (int i = 0; i < ndim; ++i) p.push_back(A.take());
  if (!A.ok) return A.err;
  const auto d = f.dimensions();
  if ((int)d.size() != ndim) return bad("ipx_halo_table: f must be scalar-valued with ndim dimensions");
  int32_t n[3] = {0, 0, 0};
  const int32_t* pp[3] = {nullptr, nullptr, nullptr};
  for (int i = 0; i < ndim; ++i) {
    n[i] = (int32_t)d[i];
    pp[i] = static_cast<const int32_t*>(data_or_null(p[i]));
  }
  int rc = f.element_type() == ffi::F64
               ? ipx_halo_table_f64(ndim, static_cast<const double*>(f.untyped_data()), n, periodic_mask, pp,
                                    static_cast<double*>(out->untyped_data()), stream)
               : ipx_halo_table_f32(ndim, static_cast<const float*>(f.untyped_data()), n, periodic_mask, pp,
                                    static_cast<float*>(out->untyped_data()), stream);
  return to_error(rc, "ipx_halo_table");
}

// ---------------------------------------------------------------------------------------------
// ipx_slopes_*: operands x, f, bc_start_value, bc_end_value (empty = none); results out, workspace
// ---------------------------------------------------------------------------------------------
ffi::Error SlopesImpl(cudaStream_t stream, ffi::AnyBuffer x, ffi::AnyBuffer f, ffi::AnyBuffer v0, ffi::AnyBuffer v1,
                      ffi::Result<ffi::AnyBuffer> out, ffi::Result<ffi::AnyBuffer> ws, int32_t method, int32_t axis,
                      double c, int32_t bc_start, int32_t bc_end) {
  const auto d = f.dimensions();
  if (axis < 0 || axis >= (int)d.size()) return bad("ipx_slopes: axis out of range");
  const int64_t outer = product(d, 0, axis), inner = product(d, axis + 1, d.size());
  ipx_slope_opts o;
  o.c = c;
  o.bc_start = bc_start;
  o.bc_end = bc_end;
  o.bc_start_value = data_or_null(v0);
  o.bc_end_value = data_or_null(v1);
  int rc = f.element_type() == ffi::F64
               ? ipx_slopes_f64(static_cast<const double*>(x.untyped_data()), (int32_t)d[axis],
                                static_cast<const double*>(f.untyped_data()), outer, inner, method, &o,
                                static_cast<double*>(out->untyped_data()), ws->untyped_data(), stream)
               : ipx_slopes_f32(static_cast<const float*>(x.untyped_data()), (int32_t)d[axis],
                                static_cast<const float*>(f.untyped_data()), outer, inner, method, &o,
                                static_cast<float*>(out->untyped_data()), ws->untyped_data(), stream);
  return to_error(rc, "ipx_slopes");
}

// ipx_slopes_vjp_*: operands x, g (cotangent of the slopes); results grad_f, workspace
ffi::Error SlopesVjpImpl(cudaStream_t stream, ffi::AnyBuffer x, ffi::AnyBuffer g, ffi::Result<ffi::AnyBuffer> out,
                         ffi::Result<ffi::AnyBuffer> ws, int32_t method, int32_t axis, double c, int32_t bc_start,
                         int32_t bc_end) {
  const auto d = g.dimensions();
  if (axis < 0 || axis >= (int)d.size()) return bad("ipx_slopes_vjp: axis out of range");
  const int64_t outer = product(d, 0, axis), inner = product(d, axis + 1, d.size());
  ipx_slope_opts o;
  o.c = c;
  o.bc_start = bc_start;
  o.bc_end = bc_end;
  o.bc_start_value = nullptr;  // boundary values are constants of the map: not part of its transpose
  o.bc_end_value = nullptr;
  int rc = g.element_type() == ffi::F64
               ? ipx_slopes_vjp_f64(static_cast<const double*>(x.untyped_data()), (int32_t)d[axis],
                                    static_cast<const double*>(g.untyped_data()), outer, inner, method, &o, 0,
                                    static_cast<double*>(out->untyped_data()), ws->untyped_data(), stream)
               : ipx_slopes_vjp_f32(static_cast<const float*>(x.untyped_data()), (int32_t)d[axis],
                                    static_cast<const float*>(g.untyped_data()), outer, inner, method, &o, 0,
                                    static_cast<float*>(out->untyped_data()), ws->untyped_data(), stream);
  return to_error(rc, "ipx_slopes_vjp");
}

// ---------------------------------------------------------------------------------------------
// ipx_slopes_pack_* (fused chain -> records) and ipx_pack_* (tables -> records)
// ---------------------------------------------------------------------------------------------
ffi::Error SlopesPackImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::Result<ffi::AnyBuffer> out, int32_t ndim,
                          int32_t method, double c, int64_t batch) {
  Args A{args};
  std::vector<ffi::AnyBuffer> k;
  for (int i = 0; i < ndim; ++i) k.push_back(A.take());
  ffi::AnyBuffer f = A.take();
  if (!A.ok) return A.err;
  const auto d = f.dimensions();
  int32_t n[3] = {0, 0, 0};
  const void* kp[3];
  for (int i = 0; i < ndim; ++i) {
    n[i] = (int32_t)d[i];
    kp[i] = k[i].untyped_data();
  }
  ipx_slope_opts o{c, 0, 0, nullptr, nullptr};
  int rc = f.element_type() == ffi::F64
               ? ipx_slopes_pack_f64(ndim, reinterpret_cast<const double* const*>(kp), n,
                                     static_cast<const double*>(f.untyped_data()), batch, method, &o,
                                     static_cast<double*>(out->untyped_data()), stream)
               : ipx_slopes_pack_f32(ndim, reinterpret_cast<const float* const*>(kp), n,
                                     static_cast<const float*>(f.untyped_data()), batch, method, &o,
                                     static_cast<float*>(out->untyped_data()), stream);
  return to_error(rc, "ipx_slopes_pack");
}

ffi::Error PackImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::Result<ffi::AnyBuffer> out, int32_t ndim) {
  Args A{args};
  const void* tp[8];
  int64_t total = 0;
  for (int i = 0; i < (1 << ndim); ++i) {
    ffi::AnyBuffer t = A.take();
    tp[i] = t.untyped_data();
    total = (int64_t)t.element_count();
  }
  if (!A.ok) return A.err;
  int rc = out->element_type() == ffi::F64
               ? ipx_pack_f64(ndim, reinterpret_cast<const double* const*>(tp), total,
                              static_cast<double*>(out->untyped_data()), stream)
               : ipx_pack_f32(ndim, reinterpret_cast<const float* const*>(tp), total,
                              static_cast<float*>(out->untyped_data()), stream);
  return to_error(rc, "ipx_pack");
}

// ---------------------------------------------------------------------------------------------
// periodic axis preparation: knots (xpad, perm, + workspace) and the materialised table pad
// ---------------------------------------------------------------------------------------------
ffi::Error PeriodicKnotsImpl(cudaStream_t stream, ffi::AnyBuffer x, ffi::Result<ffi::AnyBuffer> xpad,
                             ffi::Result<ffi::AnyBuffer> perm, ffi::Result<ffi::AnyBuffer> ws, double period) {
  const int32_t n = (int32_t)x.element_count();
  int rc = x.element_type() == ffi::F64
               ? ipx_periodic_knots_f64(static_cast<const double*>(x.untyped_data()), n, period,
                                        static_cast<double*>(xpad->untyped_data()),
                                        static_cast<int32_t*>(perm->untyped_data()), ws->untyped_data(), stream)
               : ipx_periodic_knots_f32(static_cast<const float*>(x.untyped_data()), n, period,
                                        static_cast<float*>(xpad->untyped_data()),
                                        static_cast<int32_t*>(perm->untyped_data()), ws->untyped_data(), stream);
  return to_error(rc, "ipx_periodic_knots");
}

ffi::Error PadPeriodicImpl(cudaStream_t stream, ffi::AnyBuffer in, ffi::AnyBuffer perm, ffi::Result<ffi::AnyBuffer> out,
                           int32_t axis) {
  const auto d = in.dimensions();
  if (axis < 0 || axis >= (int)d.size()) return bad("ipx_pad_periodic: axis out of range");
  const int64_t outer = product(d, 0, axis), inner = product(d, axis + 1, d.size());
  const int32_t* pp = static_cast<const int32_t*>(data_or_null(perm));
  int rc = in.element_type() == ffi::F64
               ? ipx_pad_periodic_f64(static_cast<const double*>(in.untyped_data()), outer, (int32_t)d[axis], inner, pp,
                                      static_cast<double*>(out->untyped_data()), stream)
               : ipx_pad_periodic_f32(static_cast<const float*>(in.untyped_data()), outer, (int32_t)d[axis], inner, pp,
                                      static_cast<float*>(out->untyped_data()), stream);
  return to_error(rc, "ipx_pad_periodic");
}

// ---------------------------------------------------------------------------------------------
// ipx_vjp_tables_*: operands q[ndim], knots[ndim], perm[ndim], params, cotangent; results: the
// ntab gradient tables (zeroed here: the C entry point accumulates)
// ---------------------------------------------------------------------------------------------
ffi::Error VjpTablesImpl(cudaStream_t stream, ffi::RemainingArgs args, ffi::RemainingRets rets, int32_t ndim,
                         int32_t method_class, int32_t periodic_mask, int64_t batch, int32_t ntab,
                         ffi::Span<const int64_t> n64) {
  Args A{args};
  const void* qp[3];
  const void* kp[3];
  const int32_t* pp[3];
  int32_t n[3] = {0, 0, 0};
  int64_t nq = 0;
  ffi::DataType dt = ffi::F64;
  for (int i = 0; i < ndim; ++i) {
    ffi::AnyBuffer b = A.take();
    qp[i] = b.untyped_data();
    nq = (int64_t)b.element_count();
    dt = b.element_type();
  }
  for (int i = 0; i < ndim; ++i) kp[i] = A.take().untyped_data();
  for (int i = 0; i < ndim; ++i) pp[i] = static_cast<const int32_t*>(data_or_null(A.take()));
  ffi::AnyBuffer params = A.take(), cot = A.take();
  if (!A.ok) return A.err;
  for (int i = 0; i < ndim; ++i) n[i] = (int32_t)n64[i];
  void* gp[8];
  for (int i = 0; i < ntab; ++i) {
    auto g = rets.get<ffi::AnyBuffer>(i);
    if (!g.has_value()) return g.error();
    gp[i] = (*g.value()).untyped_data();
    if (cudaMemsetAsync(gp[i], 0, (*g.value()).size_bytes(), stream) != cudaSuccess) return bad("cudaMemsetAsync failed");
  }
  const ipx_params* P = static_cast<const ipx_params*>(params.untyped_data());
  int rc = dt == ffi::F64
               ? ipx_vjp_tables_f64(ndim, reinterpret_cast<const double* const*>(qp), nq,
                                    reinterpret_cast<const double* const*>(kp), pp, n, batch, method_class,
                                    periodic_mask, P, 1, static_cast<const double*>(cot.untyped_data()),
                                    reinterpret_cast<double* const*>(gp), stream)
               : ipx_vjp_tables_f32(ndim, reinterpret_cast<const float* const*>(qp), nq,
                                    reinterpret_cast<const float* const*>(kp), pp, n, batch, method_class, periodic_mask,
                                    P, 1, static_cast<const float*>(cot.untyped_data()),
                                    reinterpret_cast<float* const*>(gp), stream);
  return to_error(rc, "ipx_vjp_tables");
}

ffi::Error BinIndexImpl(cudaStream_t stream, ffi::AnyBuffer knots, ffi::AnyBuffer q, ffi::Result<ffi::AnyBuffer> idx) {
  const int32_t n = (int32_t)knots.element_count();
  const int64_t nq = (int64_t)q.element_count();
  int rc = q.element_type() == ffi::F64
               ? ipx_bin_index_f64(static_cast<const double*>(knots.untyped_data()), n,
                                   static_cast<const double*>(q.untyped_data()), nq,
                                   static_cast<int32_t*>(idx->untyped_data()), stream)
               : ipx_bin_index_f32(static_cast<const float*>(knots.untyped_data()), n,
                                   static_cast<const float*>(q.untyped_data()), nq,
                                   static_cast<int32_t*>(idx->untyped_data()), stream);
  return to_error(rc, "ipx_bin_index");
}

}  // namespace

using Stream = ffi::PlatformStream<cudaStream_t>;

XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxEval, EvalImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .RemainingArgs()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("ndim")
                                  .Attr<int32_t>("layout")
                                  .Attr<int32_t>("method_class")
                                  .Attr<int32_t>("periodic_mask")
                                  .Attr<int64_t>("batch")
                                  .Attr<int32_t>("ntab"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxEvalStencil, StencilImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .RemainingArgs()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("ndim")
                                  .Attr<int32_t>("slope_method")
                                  .Attr<double>("c")
                                  .Attr<int32_t>("periodic_mask")
                                  .Attr<int32_t>("hint")
                                  .Attr<ffi::Span<const int64_t>>("n"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxHaloTable, HaloImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().RemainingArgs().Ret<ffi::AnyBuffer>().Attr<int32_t>("ndim").Attr<int32_t>(
                                  "periodic_mask"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxSlopes, SlopesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("method")
                                  .Attr<int32_t>("axis")
                                  .Attr<double>("c")
                                  .Attr<int32_t>("bc_start")
                                  .Attr<int32_t>("bc_end"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxSlopesVjp, SlopesVjpImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("method")
                                  .Attr<int32_t>("axis")
                                  .Attr<double>("c")
                                  .Attr<int32_t>("bc_start")
                                  .Attr<int32_t>("bc_end"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxSlopesPack, SlopesPackImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .RemainingArgs()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("ndim")
                                  .Attr<int32_t>("method")
                                  .Attr<double>("c")
                                  .Attr<int64_t>("batch"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxPack, PackImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().RemainingArgs().Ret<ffi::AnyBuffer>().Attr<int32_t>("ndim"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxPeriodicKnots, PeriodicKnotsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<double>("period"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxPadPeriodic, PadPeriodicImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Attr<int32_t>("axis"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxVjpTables, VjpTablesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<Stream>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int32_t>("ndim")
                                  .Attr<int32_t>("method_class")
                                  .Attr<int32_t>("periodic_mask")
                                  .Attr<int64_t>("batch")
                                  .Attr<int32_t>("ntab")
                                  .Attr<ffi::Span<const int64_t>>("n"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(IpxBinIndex, BinIndexImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Arg<ffi::AnyBuffer>().Arg<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>());
#endif  // IPX_HAVE_XLA_FFI
