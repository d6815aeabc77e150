// placeholder, replaced below
#include "ss_common.cuh"
using namespace ss;
extern "C" int ss_conv_wgrad_bf16(const ss_block_desc* g, const void* x, const void* g_bf16, float* g_w, void* stream) {
    set_error("ss_conv_wgrad_bf16: not built yet");
    return SS_EUNSUPPORTED;
}
