// Developer harness: compiles ONE instantiation of the single-step FAST kernel (seconds instead of minutes) so that its
// SASS can be inspected:  scripts/dev/sass_lines.sh
#include "../../crystalgrowth_b200/csrc/kob_fast.cuh"
#ifdef WITH_FAST2
#include "../../crystalgrowth_b200/csrc/kob_fast2.cuh"
void* use2() { return (void*)kob::kob_step_fast2<6, true, false>; }
#endif
void* use1() { return (void*)kob::kob_step_fast<6, 1, false>; }
