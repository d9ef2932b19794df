// kernel_fused.cu -- placeholder until the fused kernel lands (next commit).
#include "kernels.h"
namespace b200vr {
bool FusedSupported(const FusedPlan&, const char** why) { if (why) *why = "fused kernel not built yet"; return false; }
cudaError_t LaunchFused(const FusedPlan&, cudaStream_t, int* launches) { if (launches) *launches = 0; return cudaErrorNotSupported; }
}
