// mr_fast.cu -- tuned packed-RGB path (placeholder until the bit-sliced kernel lands).
#include "mr_internal.h"
bool mr_fast_supported(const MrJob&) { return false; }
cudaError_t mr_launch_fast(const MrJob&, int, cudaStream_t, int*) { return cudaErrorNotSupported; }
