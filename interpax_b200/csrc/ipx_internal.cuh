// ipx_internal.cuh -- things shared between translation units of libipx.so (not part of the ABI).
#pragma once
#include <stdint.h>

namespace ipx {

// "Run only if": a launch made while this is set hands the condition to its kernels, which return
// at once unless bit (*flag) of `accept` is set.  That is how a call chooses between kernel shapes
// and table layouts ON THE DEVICE (ipx_stencil.cu: a probe kernel classifies the query stream and
// writes *flag; every candidate is enqueued; one runs) without a host synchronisation.  Thread-local
// and consumed by the next launch_eval on this thread: no state survives a call.
struct RunIf {
  const int32_t* flag;
  uint32_t accept;
};
extern thread_local RunIf g_run_if;

// Diagnostics only: kernels enqueued by this process through the evaluation entry points
// (ipx_launch_count), so a benchmark can report how many launches its timed region made.
void count_launches(int n);

}  // namespace ipx
