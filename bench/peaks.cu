// peaks.cu (was tools/microbench.cu) — the two denominators SURVEY.md §8(d) asks for besides the HBM copy bandwidth of
// MEASURED_PEAKS.json: random-sector gather bandwidth (the access pattern of BVH nodes, triangle records,
// texels and sampler tables) and float reduction (red.global.add) throughput (the gradient scatter).
//
//   make -C zdr_b200/csrc microbench && tools/microbench > profiles/r01_microbench.json
//
// Gather: every lane reads one random, naturally aligned record of 32 / 64 / 128 bytes (2 / 4 / 8 x 16-byte
// vector loads) from a working set that is either L2-resident (16 MiB: the C3 texture) or far larger than the
// 126 MB L2 (2 GiB: the C5 tree).  Records are independent (4 in flight per lane), like different rays.
// Scatter: red.global.add.f32 / .v4.f32 to random distinct texels of a 16 MiB texture (C3), and to one texel
// (C2 / C4: every lane of every warp hits the same 16 bytes).
// Timing: CUDA events, 3 warm-up launches, best of 5; grid = 148 SMs x 16 blocks of 256 threads.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int VEC>   // VEC float4 per record
__global__ void k_gather(const float4* __restrict__ buf, uint32_t rec_mask, int iters, float* __restrict__ sink) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    uint32_t h = mix(tid * 2654435761u + 12345u);
    for (int it = 0; it < iters; it += 4) {
        float4 v[4][VEC];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            h = mix(h + 0x9e3779b9u);
            const float4* p = buf + (size_t)(h & rec_mask) * VEC;
#pragma unroll
            for (int q = 0; q < VEC; ++q) v[k][q] = __ldg(p + q);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int q = 0; q < VEC; ++q) acc += v[k][q].x + v[k][q].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int V4, int SAME>
__global__ void k_red(float4* __restrict__ tex, uint32_t texel_mask, int iters) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t h = mix(tid * 2654435761u + 777u);
    for (int it = 0; it < iters; ++it) {
        h = mix(h + 0x9e3779b9u);
        float4* p = tex + (SAME ? 0u : (h & texel_mask));
        if (V4) asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(1.0f), "f"(0.5f), "f"(0.25f), "f"(0.125f) : "memory");
        else asm volatile("red.global.add.f32 [%0], %1;" ::"l"(reinterpret_cast<float*>(p) + (h >> 30)), "f"(1.0f) : "memory");
    }
}

template <typename F>
static float best_ms(F launch) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    float best = 1e30f;
    for (int i = 0; i < 5; ++i) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int blocks = prop.multiProcessorCount * 16, threads = 256;
    const size_t big = (size_t)2 << 30, small = (size_t)16 << 20;
    float4* buf;
    float* sink;
    CK(cudaMalloc(&buf, big));
    CK(cudaMalloc(&sink, 16));
    CK(cudaMemset(buf, 0, big));
    printf("{\"device\": \"%s\", \"sms\": %d, \"l2_bytes\": %d,\n \"gather\": [", prop.name, prop.multiProcessorCount, prop.l2CacheSize);
    bool first = true;
    for (int ws = 0; ws < 2; ++ws) {
        const size_t bytes = ws == 0 ? small : big;
        for (int rec = 32; rec <= 128; rec *= 2) {
            const uint32_t mask = (uint32_t)(bytes / rec - 1);
            const int iters = ws == 0 ? 256 : 64;
            float ms = 0.0f;
            if (rec == 32) ms = best_ms([&] { k_gather<2><<<blocks, threads>>>(buf, mask, iters, sink); });
            if (rec == 64) ms = best_ms([&] { k_gather<4><<<blocks, threads>>>(buf, mask, iters, sink); });
            if (rec == 128) ms = best_ms([&] { k_gather<8><<<blocks, threads>>>(buf, mask, iters, sink); });
            const double recs = (double)blocks * threads * iters;
            printf("%s\n  {\"working_set_mib\": %zu, \"record_bytes\": %d, \"grecords_per_s\": %.2f, \"gb_per_s\": %.1f}", first ? "" : ",",
                   bytes >> 20, rec, recs / ms * 1e-6, recs * rec / ms * 1e-6);
            first = false;
        }
    }
    printf("],\n \"red\": [");
    first = true;
    for (int v4 = 0; v4 < 2; ++v4)
        for (int same = 0; same < 2; ++same) {
            const uint32_t mask = (uint32_t)(small / 16 - 1);
            const int iters = same ? 32 : 256;
            float ms = 0.0f;
            float4* tex = buf;
            if (!v4 && !same) ms = best_ms([&] { k_red<0, 0><<<blocks, threads>>>(tex, mask, iters); });
            if (!v4 && same) ms = best_ms([&] { k_red<0, 1><<<blocks, threads>>>(tex, mask, iters); });
            if (v4 && !same) ms = best_ms([&] { k_red<1, 0><<<blocks, threads>>>(tex, mask, iters); });
            if (v4 && same) ms = best_ms([&] { k_red<1, 1><<<blocks, threads>>>(tex, mask, iters); });
            const double ops = (double)blocks * threads * iters;
            printf("%s\n  {\"op\": \"red.global.add.%sf32\", \"addresses\": \"%s\", \"gops_per_s\": %.2f, \"gb_per_s_payload\": %.1f}", first ? "" : ",",
                   v4 ? "v4." : "", same ? "one texel" : "random texels of 16 MiB", ops / ms * 1e-6, ops * (v4 ? 16 : 4) / ms * 1e-6);
            first = false;
        }
    printf("]}\n");
    return 0;
}
