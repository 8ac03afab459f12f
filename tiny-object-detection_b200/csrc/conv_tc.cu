// placeholder until the tcgen05 kernel lands: no layer is taken by the tensor-core path
#include "conv_tc.h"
#include "common.h"
namespace tod {
struct ConvTc {};
bool conv_tc_supported(const ConvGeom&, int64_t, const void*, const void*) { return false; }
int conv_tc_create(const ConvTcArgs&, ConvTc**) { return fail(TOD_ERR_UNSUPPORTED, "conv_tc not built"); }
int conv_tc_launch(ConvTc*, int, cudaStream_t) { return fail(TOD_ERR_UNSUPPORTED, "conv_tc not built"); }
void conv_tc_destroy(ConvTc*) {}
}  // namespace tod
extern "C" int tod_i8_gemm_selftest(int, int, int, int, int, float*, double*) { return tod::fail(TOD_ERR_UNSUPPORTED, "conv_tc not built"); }
