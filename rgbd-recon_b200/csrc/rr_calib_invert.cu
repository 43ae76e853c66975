#include "rr_context.h"
namespace rr {
int launch_calib_invert(rr_ctx* c, int, const uint32_t*, float4*) { return fail(c, RR_ERR_UNSUPPORTED, "calib_invert: not built yet"); }
}
