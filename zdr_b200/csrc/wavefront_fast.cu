// wavefront_fast.cu — wavefront.cu compiled a second time as the "fast" build (see render_fast.cu).
#define ZDR_FAST_MATH 1
#define launch_path_wavefront launch_path_wavefront_fast
#include "wavefront.cu"
