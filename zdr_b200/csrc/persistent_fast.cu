// persistent_fast.cu — persistent.cu compiled a second time as the "fast" build (see render_fast.cu).
#define ZDR_FAST_MATH 1
#define launch_path_persistent launch_path_persistent_fast
#include "persistent.cu"
