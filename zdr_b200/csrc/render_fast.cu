// render_fast.cu — the render kernels of render.cu compiled a second time as the "fast" build (ZDR_FLAG_FAST_MATH; Makefile:
// --use_fast_math).  Only the launchers api.cu dispatches to get a second name; everything else of render.cu that has
// external linkage (probes, counters) exists once, in the exact build.
#define ZDR_FAST_MATH 1
#define launch_forward launch_forward_fast
#define launch_backward launch_backward_fast
#include "render.cu"
