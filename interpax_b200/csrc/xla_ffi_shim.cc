// xla_ffi_shim.cc -- XLA FFI handlers that forward to the C ABI of libipx.so.
//
// NOT BUILT IN THIS ENVIRONMENT: jaxlib (which ships xla/ffi/api/ffi.h, see
// jax.ffi.include_dir()) is not installed and cannot be, so this file has never been
// compiled or run here.  It is written against the documented jax.ffi / XLA FFI API
// (jax >= 0.4.38) and compiles to nothing when the header is absent.  INTEGRATION.md shows
// the Python side (registration, custom_jvp, vmap rules).
//
// Build (where JAX is available):
//   g++ -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include xla_ffi_shim.cc -L. -lipx -o libipx_xla.so
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define IPX_HAVE_XLA_FFI 1
#endif
#endif

#ifdef IPX_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include "ipx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error to_error(int rc, const char* what) {
  if (rc == IPX_OK) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, std::string(what) + " failed with code " + std::to_string(rc));
}

// interp3d, cubic family, float64.  Operands: xq yq zq (nq) | x y z knots | 8 tables |
// params (ipx_params as a uint8 buffer ON DEVICE: derivative, extrap and period are traced
// values in the reference, interpax/_spline.py:444,595,808) | optional perm buffers.
ffi::Error Eval3dF64(cudaStream_t stream, ffi::Buffer<ffi::F64> xq, ffi::Buffer<ffi::F64> yq,
                     ffi::Buffer<ffi::F64> zq, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F64> y,
                     ffi::Buffer<ffi::F64> z, ffi::Buffer<ffi::F64> f, ffi::Buffer<ffi::F64> fx,
                     ffi::Buffer<ffi::F64> fy, ffi::Buffer<ffi::F64> fz, ffi::Buffer<ffi::F64> fxy,
                     ffi::Buffer<ffi::F64> fxz, ffi::Buffer<ffi::F64> fyz, ffi::Buffer<ffi::F64> fxyz,
                     ffi::Buffer<ffi::U8> params, ffi::Buffer<ffi::S32> permx,
                     ffi::Buffer<ffi::S32> permy, ffi::Buffer<ffi::S32> permz, int32_t method_class,
                     int32_t periodic_mask, int64_t batch, ffi::ResultBuffer<ffi::F64> out) {
  const double* tabs[8] = {f.typed_data(),   fx.typed_data(),  fy.typed_data(),  fz.typed_data(),
                           fxy.typed_data(), fxz.typed_data(), fyz.typed_data(), fxyz.typed_data()};
  const auto dims = f.dimensions();
  const int64_t nq = static_cast<int64_t>(xq.element_count());
  int rc = ipx_eval3d_f64(
      xq.typed_data(), yq.typed_data(), zq.typed_data(), nq, x.typed_data(), y.typed_data(),
      z.typed_data(), (periodic_mask & 1) ? permx.typed_data() : nullptr,
      (periodic_mask & 2) ? permy.typed_data() : nullptr,
      (periodic_mask & 4) ? permz.typed_data() : nullptr, static_cast<int32_t>(dims[0]),
      static_cast<int32_t>(dims[1]), static_cast<int32_t>(dims[2]), tabs, IPX_SOA, batch, method_class,
      periodic_mask, reinterpret_cast<const ipx_params*>(params.typed_data()),
      /*params_on_device=*/1, out->typed_data(), /*idx_out=*/nullptr, stream);
  return to_error(rc, "ipx_eval3d_f64");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    IpxEval3dF64, Eval3dF64,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>()  // xq
        .Arg<ffi::Buffer<ffi::F64>>()  // yq
        .Arg<ffi::Buffer<ffi::F64>>()  // zq
        .Arg<ffi::Buffer<ffi::F64>>()  // x
        .Arg<ffi::Buffer<ffi::F64>>()  // y
        .Arg<ffi::Buffer<ffi::F64>>()  // z
        .Arg<ffi::Buffer<ffi::F64>>()  // f
        .Arg<ffi::Buffer<ffi::F64>>()  // fx
        .Arg<ffi::Buffer<ffi::F64>>()  // fy
        .Arg<ffi::Buffer<ffi::F64>>()  // fz
        .Arg<ffi::Buffer<ffi::F64>>()  // fxy
        .Arg<ffi::Buffer<ffi::F64>>()  // fxz
        .Arg<ffi::Buffer<ffi::F64>>()  // fyz
        .Arg<ffi::Buffer<ffi::F64>>()  // fxyz
        .Arg<ffi::Buffer<ffi::U8>>()   // ipx_params bytes (device)
        .Arg<ffi::Buffer<ffi::S32>>()  // permx
        .Arg<ffi::Buffer<ffi::S32>>()  // permy
        .Arg<ffi::Buffer<ffi::S32>>()  // permz
        .Attr<int32_t>("method_class")
        .Attr<int32_t>("periodic_mask")
        .Attr<int64_t>("batch")
        .Ret<ffi::Buffer<ffi::F64>>());
// The 1-D / 2-D / float32 handlers and the ipx_slopes_* / ipx_periodic_knots_* handlers follow
// the same pattern (one handler per C-ABI symbol of include/ipx.h).
#endif  // IPX_HAVE_XLA_FFI
