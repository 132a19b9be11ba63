// placeholder, replaced below
#ifndef KOB_FAST_CUH
#define KOB_FAST_CUH
#include "kob_common.cuh"
namespace kob {
template <typename real>
int launch_step_fast(const StepArgs<real>&, bool, cudaStream_t) { return -6; }
}
#endif
