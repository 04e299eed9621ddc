// raster.cu — TEMPORARY stubs (rasteriser under construction): every entry fails loudly.
#include "common.cuh"
namespace ogrecu { int rt_release(ogrecu_rt*) { return 0; } }
#define STUB(name, ...) int name(__VA_ARGS__) { ogrecu::set_error(#name ": rasteriser not built yet"); return OGRECU_ERR_UNSUPPORTED; }
extern "C" {
STUB(ogrecu_tex_create, ogrecu_ctx*, uint32_t, uint32_t, ogrecu_tex**)
STUB(ogrecu_tex_destroy, ogrecu_tex*)
STUB(ogrecu_tex_upload, ogrecu_tex*, const uint8_t*, size_t)
STUB(ogrecu_tex_download, ogrecu_tex*, uint8_t*, size_t)
STUB(ogrecu_tex_device_ptr, ogrecu_tex*, void**)
STUB(ogrecu_rt_create, ogrecu_ctx*, uint32_t, uint32_t, ogrecu_rt**)
STUB(ogrecu_rt_destroy, ogrecu_rt*)
STUB(ogrecu_rt_size, const ogrecu_rt*, uint32_t*, uint32_t*)
STUB(ogrecu_rt_clear, ogrecu_rt*, unsigned, const float*, float)
STUB(ogrecu_draw, ogrecu_rt*, const ogrecu_draw_desc*)
STUB(ogrecu_flush, ogrecu_rt*)
STUB(ogrecu_draw_views, ogrecu_rt* const*, uint32_t, const ogrecu_draw_desc*, const float*, const float*, const float*)
STUB(ogrecu_rt_set_strip, ogrecu_rt*, uint32_t, uint32_t)
STUB(ogrecu_rt_read_colour, ogrecu_rt*, uint32_t, uint32_t, uint32_t, uint32_t, uint8_t*, size_t)
STUB(ogrecu_rt_read_depth, ogrecu_rt*, uint32_t, uint32_t, uint32_t, uint32_t, float*, size_t)
STUB(ogrecu_rt_write_colour, ogrecu_rt*, const uint8_t*, size_t)
STUB(ogrecu_rt_write_depth, ogrecu_rt*, const float*, size_t)
STUB(ogrecu_rt_colour_device_ptr, ogrecu_rt*, void**)
STUB(ogrecu_rt_depth_device_ptr, ogrecu_rt*, void**)
STUB(ogrecu_rt_last_stats, ogrecu_rt*, ogrecu_raster_stats*)
}
