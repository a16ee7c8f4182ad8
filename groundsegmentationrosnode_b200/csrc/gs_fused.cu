// gs_fused.cu — placeholder: the fused single-kernel pipeline is not built yet; every call routes to the staged kernels.
#include "gs_internal.h"

namespace gs {
int fused_configure(FusedState& fs, const KParams& kp, const Luts& lut, const SgTaps& sg, bool smooth, bool repair, int) {
    fs.kp = kp; fs.lut = lut; fs.sg = sg; fs.smooth = smooth; fs.repair = repair; fs.configured = false;
    return 0;
}
bool fused_supported(const FusedState& fs) { return fs.configured && !fs.disabled; }
int fused_run(FusedState&, const Launch&, const float4*, const long long*, int, uint16_t*, uint16_t*, uint16_t*, float4*, unsigned*, Counters*, bool) {
    return cudaErrorNotSupported;
}
void fused_release(FusedState& fs) { fs.configured = false; }
}  // namespace gs
