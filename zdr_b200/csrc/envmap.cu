// envmap.cu — environment-map import: luminance*sin(theta) weight map on the GPU, MIS
// compensation and alias tables on the host (the reference does the same split:
// generate_weight_map_kernel envmap.py:136-159 on the device, envmap.py:167-194 in host Python).
#include <cmath>
#include <cstring>
#include <vector>

#include "envmap.cuh"
#include "shading.cuh"

namespace zdr {

namespace {

constexpr int kTap = 8;                        // envmap.py:144: taps -8..8
__constant__ float c_gauss[(2 * kTap + 1) * (2 * kTap + 1)];

// envmap.py:136-158: one thread per cell of the 512x256 sampling map
__global__ void k_weight_map(const SceneView sc, float* __restrict__ scale_map) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= kEnvMapW || py >= kEnvMapH) return;
    const float cx = (float)px + 0.5f, cy = (float)py + 0.5f;
    float sum_weight = 0.0f, sum_scale = 0.0f;
    const float filter_step = 0.125f;
    for (int dy = -kTap; dy <= kTap; ++dy)
        for (int dx = -kTap; dx <= kTap; ++dx) {
            const float ox = (float)dx * filter_step, oy = (float)dy * filter_step;
            const float2 uv = make_float2((cx + ox) / (float)kEnvMapW, (cy + oy) / (float)kEnvMapH);
            const float3 c = env_texture(sc, uv);
            const float scale = 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z;
            float sn, cs;
            sincos_contract(uv.y * kPi, &sn, &cs);
            const float weight = c_gauss[(dy + kTap) * (2 * kTap + 1) + (dx + kTap)];
            const float value = weight * fminf(scale * sn, 1e8f);
            sum_weight += weight;
            sum_scale += value;
        }
    scale_map[py * kEnvMapW + px] = sum_scale / sum_weight;
}

// envmap.py:17-57 — Vose alias method in float64, stack discipline of the Python lists
void create_alias_table(const double* values, int n, float* prob_out, int* alias_out, double* pdf_out) {
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += std::fabs(values[i]);
    for (int i = 0; i < n; ++i) pdf_out[i] = sum == 0.0 ? 1.0 / n : std::fabs(values[i]) / sum;
    const double ratio = sum > 0.0 ? n / sum : 1.0;
    std::vector<double> prob(n);
    std::vector<int> over, under;
    over.reserve(n); under.reserve(n);
    for (int i = 0; i < n; ++i) {
        prob[i] = values[i] * ratio;
        alias_out[i] = i;
        if (prob[i] > 1.0) over.push_back(i);
        else if (prob[i] < 1.0) under.push_back(i);
    }
    while (!over.empty() && !under.empty()) {
        const int o = over.back(), u = under.back();
        over.pop_back(); under.pop_back();
        prob[o] -= 1.0 - prob[u];
        alias_out[u] = o;
        if (prob[o] > 1.0) over.push_back(o);
        else if (prob[o] < 1.0) under.push_back(o);
    }
    for (int i : over) { prob[i] = 1.0; alias_out[i] = i; }
    for (int i : under) { prob[i] = 1.0; alias_out[i] = i; }
    for (int i = 0; i < n; ++i) prob_out[i] = (float)prob[i];
}

}  // namespace

void free_envmap(EnvMap* e) {
    if (!e) return;
    cudaFree(e->tex); cudaFree(e->alias); cudaFree(e->pdf);
    memset(e, 0, sizeof(*e));
}

#define ENV_CK(x) do { err = (x); if (err != cudaSuccess) goto done; } while (0)

cudaError_t build_envmap(const float* rgba, int w, int h_in, bool compensate_mis, cudaStream_t stream, EnvMap* out) {
    cudaError_t err = cudaSuccess;
    const int W = kEnvMapW, H = kEnvMapH, pixel_count = W * H;
    const int h = (h_in == w) ? h_in : 2 * h_in;      // envmap.py:123-125: rows duplicated for 2:1 images
    float* d_scale = nullptr;
    std::vector<float> scale_map(pixel_count);
    std::vector<float> tex((size_t)4 * w * h);
    free_envmap(out);
    for (int y = 0; y < h; ++y)
        memcpy(tex.data() + (size_t)4 * y * w, rgba + (size_t)4 * ((h == h_in) ? y : y / 2) * w, sizeof(float) * 4 * w);
    {
        float gauss[(2 * kTap + 1) * (2 * kTap + 1)];
        for (int dy = -kTap; dy <= kTap; ++dy)
            for (int dx = -kTap; dx <= kTap; ++dx) {
                const float ox = (float)dx * 0.125f, oy = (float)dy * 0.125f;
                gauss[(dy + kTap) * (2 * kTap + 1) + (dx + kTap)] = expf(-4.0f * fmaf(oy, oy, ox * ox));   // envmap.py:153
            }
        ENV_CK(cudaMemcpyToSymbolAsync(c_gauss, gauss, sizeof(gauss), 0, cudaMemcpyHostToDevice, stream));
        ENV_CK(cudaMalloc(&out->tex, sizeof(float4) * (size_t)w * h));
        ENV_CK(cudaMalloc(&out->alias, sizeof(float2) * (size_t)(H + pixel_count)));
        ENV_CK(cudaMalloc(&out->pdf, sizeof(float) * (size_t)pixel_count));
        ENV_CK(cudaMalloc(&d_scale, sizeof(float) * (size_t)pixel_count));
        ENV_CK(cudaMemcpyAsync(out->tex, tex.data(), sizeof(float4) * (size_t)w * h, cudaMemcpyHostToDevice, stream));
        out->w = w; out->h = h;
        SceneView sc;
        memset(&sc, 0, sizeof(sc));
        sc.env_tex = out->tex; sc.env_w = w; sc.env_h = h;
        k_weight_map<<<dim3(W / 16, H / 16), dim3(16, 16), 0, stream>>>(sc, d_scale);
        ENV_CK(cudaGetLastError());
        ENV_CK(cudaMemcpyAsync(scale_map.data(), d_scale, sizeof(float) * pixel_count, cudaMemcpyDeviceToHost, stream));
        ENV_CK(cudaStreamSynchronize(stream));
        if (compensate_mis) {                                                    // envmap.py:167-175
            double mean = 0.0;
            for (int i = 0; i < pixel_count; ++i) mean += scale_map[i];
            const float average_scale = (float)(mean / pixel_count);
            double wavg = 0.0;
            for (int y = 0; y < H; ++y) wavg += std::sin((y + 0.5) / H * M_PI);
            wavg /= H;
            for (int y = 0; y < H; ++y) {
                const float sub = (float)((double)average_scale * std::sin((y + 0.5) / H * M_PI) / wavg);
                for (int x = 0; x < W; ++x) {
                    const float v = scale_map[y * W + x] - sub;
                    scale_map[y * W + x] = v > 0.0f ? v : 0.0f;
                }
            }
        }
        std::vector<float> prob(H + pixel_count);                                 // envmap.py:177-194
        std::vector<int> alias(H + pixel_count);
        std::vector<double> pdfs(pixel_count), row(W), row_avg(H), pdf_y(H);
        for (int y = 0; y < H; ++y) {
            double md = 0.0;
            for (int x = 0; x < W; ++x) { row[x] = scale_map[y * W + x]; md += row[x]; }
            row_avg[y] = (float)(md / W);
            create_alias_table(row.data(), W, prob.data() + H + y * W, alias.data() + H + y * W, pdfs.data() + y * W);
        }
        create_alias_table(row_avg.data(), H, prob.data(), alias.data(), pdf_y.data());
        std::vector<float2> packed(H + pixel_count);
        std::vector<float> pdf(pixel_count);
        for (int i = 0; i < H + pixel_count; ++i) packed[i] = make_float2(prob[i], __builtin_bit_cast(float, alias[i]));
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) pdf[y * W + x] = (float)(pdfs[y * W + x] * pdf_y[y] * pixel_count);
        ENV_CK(cudaMemcpy(out->alias, packed.data(), sizeof(float2) * packed.size(), cudaMemcpyHostToDevice));
        ENV_CK(cudaMemcpy(out->pdf, pdf.data(), sizeof(float) * pdf.size(), cudaMemcpyHostToDevice));
        out->valid = true;
    }
done:
    cudaFree(d_scale);
    if (err != cudaSuccess) free_envmap(out);
    return err;
}

cudaError_t read_envmap_tables(const EnvMap* e, float* prob, int* alias, float* pdf) {
    if (!e || !e->valid) return cudaErrorInvalidValue;
    const int n = kEnvMapH + kEnvMapW * kEnvMapH;
    std::vector<float2> packed(n);
    cudaError_t err = cudaMemcpy(packed.data(), e->alias, sizeof(float2) * n, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) return err;
    for (int i = 0; i < n; ++i) { prob[i] = packed[i].x; alias[i] = __builtin_bit_cast(int, packed[i].y); }
    return cudaMemcpy(pdf, e->pdf, sizeof(float) * kEnvMapW * kEnvMapH, cudaMemcpyDeviceToHost);
}

}  // namespace zdr
