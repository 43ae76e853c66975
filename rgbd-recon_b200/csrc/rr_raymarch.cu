#include "rr_context.h"
namespace rr {
int launch_raymarch(rr_ctx* c, const rr_view*) { return fail(c, RR_ERR_UNSUPPORTED, "raymarch: not built yet"); }
}
