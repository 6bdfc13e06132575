// svoxb_build.cu -- one-shot octree build from points (placeholder until the sort-based builder lands).
#include "svoxb_common.cuh"
using namespace svoxb;
extern "C" size_t svoxb_build_work_bytes(int64_t, int32_t) { return 0; }
extern "C" int svoxb_build_octree_count(const float*, int64_t, int32_t, const float*, const float*, void*, int64_t*, void*) {
    set_error("svoxb_build_octree: not implemented yet");
    return SVOXB_EUNSUPPORTED;
}
extern "C" int svoxb_build_octree_emit(int64_t, int32_t, const void*, int64_t, int32_t*, int32_t*, int32_t*, void*) {
    set_error("svoxb_build_octree: not implemented yet");
    return SVOXB_EUNSUPPORTED;
}
