// hj_tma.cu -- TMA backend (placeholder until the plane-ring kernel lands in this file).
#include "hj_internal.h"
#include <cstdio>
struct HjTmaPlan { int dummy; };
HjTmaPlan* hj_tma_plan_create(const KGrid&, int, int, double* const*, int, char* err, int errlen) {
  snprintf(err, errlen, "TMA backend not built yet");
  return nullptr;
}
void hj_tma_plan_destroy(HjTmaPlan* p) { delete p; }
cudaError_t hj_launch_stage_tma(HjTmaPlan*, int, int, const KGrid&, const KSys&, const KStage&, int, cudaStream_t) {
  return cudaErrorNotSupported;
}
