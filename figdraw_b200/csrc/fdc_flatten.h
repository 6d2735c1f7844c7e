// Internal: native scene flattening shared by fdc_flatten.cu (pure host) and fdc_context.cu (fdc_render_frame).
#pragma once
#include <vector>

#include "../../include/figdraw_cuda.h"

namespace fdc {

// Writes the body of renderFrame (figrender.nim:1960-2002) into out[0..cap); *n_out = records needed (may exceed cap:
// the excess was not stored).  Returns nullptr or a static error message.
const char* flatten_renders(const fdc_scene& scene, const fdc_flatten_env& env, fdc_call* out, size_t cap, size_t* n_out);

}  // namespace fdc
