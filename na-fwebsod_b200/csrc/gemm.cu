// placeholder: tcgen05 FC GEMMs land here
#include "common.cuh"
using namespace nawsod;
extern "C" int64_t nawsod_fc_workspace_bytes(void) { return 0; }
extern "C" int nawsod_fc_fwd(const void*, const void*, const float*, const uint8_t*, int, int, int, int, void*, int, int,
                             void*, void*) { set_error("fc_fwd: not built"); return NAWSOD_ERR_UNSUPPORTED; }
extern "C" int nawsod_fc_bwd_x(const void*, const void*, const void*, const uint8_t*, int, int, int, int, void*, int,
                               int, void*, void*) { set_error("fc_bwd_x: not built"); return NAWSOD_ERR_UNSUPPORTED; }
extern "C" int nawsod_fc_bwd_w(const void*, const void*, int, int, int, int, float*, float*, int, void*, void*) {
  set_error("fc_bwd_w: not built"); return NAWSOD_ERR_UNSUPPORTED; }
