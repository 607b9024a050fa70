// rtfs_wavefront.cu — wavefront (ray-queue) variant of the render path.  Placeholder until the
// megakernel is measured; see DESIGN.md.
#include "rtfs_device.h"

namespace rtfs {
int render_wavefront(RtScene *, const RtCamera *, int32_t, int32_t, const RtRenderOpts *, uint8_t *, int32_t *, RtStats *) {
    return fail(RT_ERR_UNSUPPORTED, "rt_render: RT_MODE_WAVEFRONT is not built yet");
}
} // namespace rtfs
