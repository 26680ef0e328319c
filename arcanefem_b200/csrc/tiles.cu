// Tiled-gather assembly (inspector/executor) -- placeholder until the plan builder lands.
#include "afb_internal.h"

namespace afb {

int build_tile_plan(afb_ctx* ctx)
{
  (void)ctx;
  set_error("AFB_VARIANT_TILED_GATHER is not available in this build");
  return AFB_ERR_UNSUPPORTED;
}

int assemble_tiled(afb_ctx* ctx, int, const double*, int, int)
{
  (void)ctx;
  set_error("AFB_VARIANT_TILED_GATHER is not available in this build");
  return AFB_ERR_UNSUPPORTED;
}

} // namespace afb
