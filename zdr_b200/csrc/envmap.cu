// envmap.cu — environment-map import, entirely on the GPU: luminance*sin(theta) weight map (generate_weight_map_kernel,
// envmap.py:136-159), MIS compensation (envmap.py:167-175) and the 256 conditional + 1 marginal Vose alias tables
// (envmap.py:17-57,183-194, which the reference builds in interpreted Python).  One thread per table keeps the reference's
// pairing order, so the tables are identical to the oracle's, entry for entry.
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

// envmap.py:17-57 — Vose alias method in float64, with the stack discipline of the reference's Python lists (the pairing
// order decides the alias indices).  One thread builds one table; `prob`, `over`, `under` are n-entry scratch arrays.
__device__ void create_alias_table(const double* values, int n, float2* table_out, double* pdf_out, double* prob, int* over, int* under,
                                   int* alias) {
    double sum = 0.0;
    for (int i = 0; i < n; ++i) sum += fabs(values[i]);
    for (int i = 0; i < n; ++i) pdf_out[i] = sum == 0.0 ? 1.0 / n : fabs(values[i]) / sum;
    const double ratio = sum > 0.0 ? n / sum : 1.0;
    int n_over = 0, n_under = 0;
    for (int i = 0; i < n; ++i) {
        prob[i] = values[i] * ratio;
        alias[i] = i;
        if (prob[i] > 1.0) over[n_over++] = i;
        else if (prob[i] < 1.0) under[n_under++] = i;
    }
    while (n_over > 0 && n_under > 0) {
        const int o = over[--n_over], u = under[--n_under];
        prob[o] -= 1.0 - prob[u];
        alias[u] = o;
        if (prob[o] > 1.0) over[n_over++] = o;
        else if (prob[o] < 1.0) under[n_under++] = o;
    }
    for (int k = 0; k < n_over; ++k) { prob[over[k]] = 1.0; alias[over[k]] = over[k]; }
    for (int k = 0; k < n_under; ++k) { prob[under[k]] = 1.0; alias[under[k]] = under[k]; }
    for (int i = 0; i < n; ++i) table_out[i] = make_float2((float)prob[i], __int_as_float(alias[i]));
}

// sin((y + 0.5) / H * pi) of the rows of the sampling map and their mean (envmap.py:168-171): constants of the map size,
// evaluated once on the host so that the device tables do not depend on a second libm
__constant__ double c_row_weight[kEnvMapH];
__constant__ double c_row_weight_avg;

struct EnvScratch {        // device scratch of one import
    float* scale_map;      // W*H
    double* row_values;    // W*H: the (compensated) rows as binary64, input of the row tables
    double* pdfs;          // W*H
    double* prob;          // W*H
    int* over; int* under; int* alias;   // W*H each
    double* row_avg;       // H
    double* pdf_y;         // H
    float* average_scale;  // 1
};

// envmap.py:170: scale_map.mean(), accumulated in index order (one thread: 131 072 dependent additions, < 1 ms, once per import)
__global__ void k_env_mean(EnvScratch s) {
    double mean = 0.0;
    for (int i = 0; i < kEnvMapW * kEnvMapH; ++i) mean += s.scale_map[i];
    *s.average_scale = (float)(mean / (kEnvMapW * kEnvMapH));
}
// envmap.py:172-175 (MIS compensation of one row) and :185-190 (its conditional alias table p(x|y)): one thread per row
__global__ void k_env_rows(EnvScratch s, int compensate, float2* table) {
    const int y = blockIdx.x * blockDim.x + threadIdx.x;
    if (y >= kEnvMapH) return;
    const int W = kEnvMapW;
    float* row = s.scale_map + y * W;
    if (compensate) {
        const float sub = (float)((double)*s.average_scale * c_row_weight[y] / c_row_weight_avg);
        for (int x = 0; x < W; ++x) {
            const float v = row[x] - sub;
            row[x] = v > 0.0f ? v : 0.0f;
        }
    }
    double md = 0.0;
    double* vals = s.row_values + y * W;
    for (int x = 0; x < W; ++x) { vals[x] = row[x]; md += vals[x]; }
    s.row_avg[y] = (float)(md / W);
    create_alias_table(vals, W, table + kEnvMapH + y * W, s.pdfs + y * W, s.prob + y * W, s.over + y * W, s.under + y * W, s.alias + y * W);
}
// envmap.py:191-192: the marginal table p(y) over the row averages
__global__ void k_env_marginal(EnvScratch s, float2* table) {
    create_alias_table(s.row_avg, kEnvMapH, table, s.pdf_y, s.prob, s.over, s.under, s.alias);
}
// envmap.py:193-195
__global__ void k_env_pdf(EnvScratch s, float* __restrict__ pdf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kEnvMapW * kEnvMapH) return;
    pdf[i] = (float)(s.pdfs[i] * s.pdf_y[i / kEnvMapW] * (kEnvMapW * kEnvMapH));
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
    double* d_f64 = nullptr;
    int* d_i32 = nullptr;
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
        // MIS compensation and the 256 conditional + 1 marginal alias tables, all on the device and stream-ordered (the
        // reference builds them in interpreted Python, envmap.py:167-194 with TODOs at :181,196; round 1 built them in host C++)
        {
            static bool weights_set = false;
            double rw[kEnvMapH], wavg = 0.0;
            for (int y = 0; y < H; ++y) { rw[y] = std::sin((y + 0.5) / H * M_PI); wavg += rw[y]; }
            wavg /= H;
            ENV_CK(cudaMemcpyToSymbolAsync(c_row_weight, rw, sizeof(rw), 0, cudaMemcpyHostToDevice, stream));
            ENV_CK(cudaMemcpyToSymbolAsync(c_row_weight_avg, &wavg, sizeof(wavg), 0, cudaMemcpyHostToDevice, stream));
            (void)weights_set;
        }
        EnvScratch sx;
        memset(&sx, 0, sizeof(sx));
        sx.scale_map = d_scale;
        ENV_CK(cudaMalloc(&d_f64, sizeof(double) * ((size_t)3 * pixel_count + 2 * H)));
        ENV_CK(cudaMalloc(&d_i32, sizeof(int) * (size_t)3 * pixel_count + sizeof(float) * 4));
        sx.row_values = d_f64; sx.pdfs = d_f64 + pixel_count; sx.prob = d_f64 + 2 * (size_t)pixel_count;
        sx.row_avg = d_f64 + 3 * (size_t)pixel_count; sx.pdf_y = sx.row_avg + H;
        sx.over = d_i32; sx.under = d_i32 + pixel_count; sx.alias = d_i32 + 2 * (size_t)pixel_count;
        sx.average_scale = reinterpret_cast<float*>(d_i32 + 3 * (size_t)pixel_count);
        if (compensate_mis) k_env_mean<<<1, 1, 0, stream>>>(sx);
        k_env_rows<<<(H + 31) / 32, 32, 0, stream>>>(sx, compensate_mis ? 1 : 0, out->alias);
        k_env_marginal<<<1, 1, 0, stream>>>(sx, out->alias);
        k_env_pdf<<<(pixel_count + 255) / 256, 256, 0, stream>>>(sx, out->pdf);
        ENV_CK(cudaGetLastError());
        ENV_CK(cudaStreamSynchronize(stream));        // the scratch is freed below
        out->valid = true;
    }
done:
    cudaFree(d_scale); cudaFree(d_f64); cudaFree(d_i32);
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
