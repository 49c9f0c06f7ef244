// placeholder until the tcgen05 kernels land
#include "common.cuh"
extern "C" long long egm_conv2d_tc_workspace_bytes(int, int, int, int, int, int, int) { return 0; }
extern "C" int egm_conv2d_tc_supported(int, int, int, int, int, int) { return 0; }
extern "C" int egm_pack_conv_weight_tc(const float*, void*, void*, int, int, int, int, void*) { egm_set_error("tc path not built"); return EGM_E_ARCH; }
extern "C" int egm_conv2d_tc(const void*, const void*, const float*, void*, int, int, int, int, int, int, int, int, void*) { egm_set_error("tc path not built"); return EGM_E_ARCH; }
extern "C" int egm_conv2d_wgrad_tc(const void*, const void*, float*, int, int, int, int, int, int, int, int, void*) { egm_set_error("tc path not built"); return EGM_E_ARCH; }
