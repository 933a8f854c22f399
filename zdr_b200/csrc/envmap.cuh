// envmap.cuh — interface of the environment-map importance tables (envmap.cu).
#pragma once
#include <cuda_runtime.h>

namespace zdr {

struct EnvMap {
    bool valid;
    float4* tex;       // rgba, h rows of w texels (after the row duplication of envmap.py:123-125)
    float2* alias;     // [256 + 256*512] (prob, alias index bits): marginal rows then conditionals
    float* pdf;        // [256*512], pdf * pixel_count (envmap.py:197-200)
    int w, h;
};

// Scene.add_envmap -> load_envmap (envmap.py:116-202). rgba_host: h_in rows of w RGBA float texels.
cudaError_t build_envmap(const float* rgba_host, int w, int h_in, bool compensate_mis, cudaStream_t stream, EnvMap* out);
void free_envmap(EnvMap* e);
// read-back of the tables for the parity tests: prob/alias [256+131072], pdf [131072]
cudaError_t read_envmap_tables(const EnvMap* e, float* prob, int* alias, float* pdf);

}  // namespace zdr
