// Fused warp-per-world step -- placeholder until the kernel lands: reports
// "unsupported" so arb_step runs the four lane-per-world phase kernels.
#include "arb_internal.h"

bool arb_fused_supported(const arb_batch*) { return false; }
int arb_fused_step(arb_batch* b, const double* dts, int nsteps) { return arb_step_phases(b, dts, nsteps); }
void arb_fused_release(arb_batch*) {}
