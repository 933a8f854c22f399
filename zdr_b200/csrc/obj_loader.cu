// obj_loader.cu — native Wavefront OBJ ingest (host code).
//
// The reference parses OBJ files line by line in interpreted Python and builds a list of per-vertex tuples
// (load_obj.py:26-56, render.py:97): minutes for a mesh of millions of triangles.  This is the same algorithm with
// the same observable result — zdr_b200/load_obj.py is the readable statement of it and tests/test_host.py holds the
// two to identical bytes:
//   * a vertex is a distinct (position, uv, normal) triple, numbered in first-use order (load_obj.py:47-55); values
//     compare as the doubles Python parsed (so -0.0 merges with 0.0 and a missing normal, NaN, never merges);
//   * a missing vt is (0, 0), a missing vn is NaN (load_obj.py:48-49); indices are 1-based, and like a Python list
//     a non-positive index counts from the end;
//   * polygons are fan-triangulated (f0, f[i-1], f[i]) in face order (load_obj.py:63-68);
//   * if the first vertex has no normal, area-weighted vertex normals are recomputed (render.py:101-103,
//     recompute_normal.py:4-51) in float32, accumulating corner by corner in triangle order.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/zdr_b200.h"

namespace {

struct Key {
    double v[8];
    bool operator==(const Key& o) const {
        for (int i = 0; i < 8; ++i)
            if (!(v[i] == o.v[i])) return false;       // IEEE ==: NaN differs from everything, -0.0 equals 0.0
        return true;
    }
};
struct KeyHash {
    size_t operator()(const Key& k) const {
        uint64_t h = 1469598103934665603ull;
        for (int i = 0; i < 8; ++i) {
            double d = k.v[i] == 0.0 ? 0.0 : k.v[i];   // one hash for both zeros
            uint64_t b;
            memcpy(&b, &d, 8);
            h = (h ^ b) * 1099511628211ull;
            h ^= h >> 29;
        }
        return (size_t)h;
    }
};

inline const char* skip_ws(const char* p, const char* end) {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
    return p;
}
inline const char* token_end(const char* p, const char* end) {
    while (p < end && !(*p == ' ' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
    return p;
}

thread_local std::string g_obj_err;

bool parse_doubles(const char*& p, const char* end, int want, double* out, int line) {
    for (int k = 0; k < want; ++k) {
        p = skip_ws(p, end);
        const char* te = token_end(p, end);
        if (te == p) { g_obj_err = "line " + std::to_string(line) + ": too few numbers"; return false; }
        char* stop = nullptr;                      // the buffer is NUL-terminated and strtod stops at white space
        out[k] = strtod(p, &stop);
        if (stop != te) { g_obj_err = "line " + std::to_string(line) + ": bad number '" + std::string(p, te) + "'"; return false; }
        p = te;
    }
    return true;
}

// Python list indexing of `one_based - 1`
bool resolve(long one_based, size_t size, size_t* out) {
    long i = one_based - 1;
    if (i < 0) i += (long)size;
    if (i < 0 || (size_t)i >= size) return false;
    *out = (size_t)i;
    return true;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) const char* zdr_obj_last_error(void) { return g_obj_err.c_str(); }

extern "C" __attribute__((visibility("default"))) void zdr_obj_free(void* p) { free(p); }

extern "C" __attribute__((visibility("default"))) int zdr_obj_load(const char* path, float** verts_out, int* nv_out,
                                                                   int32_t** tris_out, int* nt_out) {
    if (!path || !verts_out || !nv_out || !tris_out || !nt_out) { g_obj_err = "null argument"; return ZDR_ERR_ARG; }
    *verts_out = nullptr; *tris_out = nullptr; *nv_out = 0; *nt_out = 0;
    FILE* f = fopen(path, "rb");
    if (!f) { g_obj_err = std::string("cannot open ") + path; return ZDR_ERR_ARG; }
    fseek(f, 0, SEEK_END);
    const long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)(size > 0 ? size : 0) + 1, '\0');
    if (size > 0 && fread(buf.data(), 1, (size_t)size, f) != (size_t)size) { fclose(f); g_obj_err = "read error"; return ZDR_ERR_ARG; }
    fclose(f);

    std::vector<double> pos, tex, nrm;          // 3 / 2 / 3 doubles per entry
    std::vector<Key> vertices;
    std::vector<int> table(1 << 16, -1);        // open addressing over `vertices` (first-use numbering)
    size_t table_used = 0;
    KeyHash hasher;
    std::vector<int32_t> tris;
    std::vector<int> poly;
    const double nan = std::nan("");
    const char* p = buf.data();
    const char* const end = p + buf.size() - 1;
    int line = 0;
    while (p < end) {
        ++line;
        const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!le) le = end;
        const char* q = skip_ws(p, le);
        const char* te = token_end(q, le);
        const size_t tl = (size_t)(te - q);
        if (tl == 1 && q[0] == 'v') {
            double d[3];
            q = te;
            if (!parse_doubles(q, le, 3, d, line)) return ZDR_ERR_ARG;
            pos.insert(pos.end(), d, d + 3);
        } else if (tl == 2 && q[0] == 'v' && q[1] == 't') {
            double d[2];
            q = te;
            if (!parse_doubles(q, le, 2, d, line)) return ZDR_ERR_ARG;
            tex.insert(tex.end(), d, d + 2);
        } else if (tl == 2 && q[0] == 'v' && q[1] == 'n') {
            double d[3];
            q = te;
            if (!parse_doubles(q, le, 3, d, line)) return ZDR_ERR_ARG;
            nrm.insert(nrm.end(), d, d + 3);
        } else if (tl == 1 && q[0] == 'f') {
            poly.clear();
            q = te;
            for (;;) {
                q = skip_ws(q, le);
                const char* ce = token_end(q, le);
                if (ce == q) break;
                // corner = a[/b[/c]]
                long idx[3] = {0, 0, 0};
                bool have[3] = {false, false, false};
                const char* s = q;
                for (int k = 0; k < 3 && s <= ce; ++k) {
                    const char* slash = (const char*)memchr(s, '/', (size_t)(ce - s));
                    const char* fe = slash ? slash : ce;
                    if (fe > s) {
                        char* stop = nullptr;
                        idx[k] = strtol(s, &stop, 10);
                        if (stop != fe) { g_obj_err = "line " + std::to_string(line) + ": bad index '" + std::string(s, fe) + "'"; return ZDR_ERR_ARG; }
                        have[k] = true;
                    }
                    if (!slash) break;
                    s = slash + 1;
                }
                if (!have[0]) { g_obj_err = "line " + std::to_string(line) + ": face corner without a position index"; return ZDR_ERR_ARG; }
                Key key;
                size_t r;
                if (!resolve(idx[0], pos.size() / 3, &r)) { g_obj_err = "line " + std::to_string(line) + ": position index out of range"; return ZDR_ERR_ARG; }
                memcpy(key.v, &pos[3 * r], sizeof(double) * 3);
                if (have[1]) {
                    if (!resolve(idx[1], tex.size() / 2, &r)) { g_obj_err = "line " + std::to_string(line) + ": vt index out of range"; return ZDR_ERR_ARG; }
                    key.v[3] = tex[2 * r]; key.v[4] = tex[2 * r + 1];
                } else {
                    key.v[3] = 0.0; key.v[4] = 0.0;
                }
                if (have[2]) {
                    if (!resolve(idx[2], nrm.size() / 3, &r)) { g_obj_err = "line " + std::to_string(line) + ": vn index out of range"; return ZDR_ERR_ARG; }
                    memcpy(key.v + 5, &nrm[3 * r], sizeof(double) * 3);
                } else {
                    key.v[5] = key.v[6] = key.v[7] = nan;
                }
                int vid = -1;
                bool has_nan = false;
                for (int k = 0; k < 8; ++k) has_nan = has_nan || key.v[k] != key.v[k];
                if (has_nan) {                          // equal to nothing, itself included: always a new vertex
                    vid = (int)vertices.size();
                    vertices.push_back(key);
                } else {
                    if (2 * (table_used + 1) > table.size()) {
                        std::vector<int> bigger(table.size() * 2, -1);
                        for (int v : table)
                            if (v >= 0) {
                                size_t h = hasher(vertices[v]) & (bigger.size() - 1);
                                while (bigger[h] >= 0) h = (h + 1) & (bigger.size() - 1);
                                bigger[h] = v;
                            }
                        table.swap(bigger);
                    }
                    size_t h = hasher(key) & (table.size() - 1);
                    while (table[h] >= 0 && !(vertices[table[h]] == key)) h = (h + 1) & (table.size() - 1);
                    if (table[h] < 0) {
                        table[h] = vid = (int)vertices.size();
                        vertices.push_back(key);
                        ++table_used;
                    } else {
                        vid = table[h];
                    }
                }
                poly.push_back(vid);
                q = ce;
            }
            for (size_t k = 2; k < poly.size(); ++k) { tris.push_back(poly[0]); tris.push_back(poly[k - 1]); tris.push_back(poly[k]); }
        }
        p = le < end ? le + 1 : end;
    }
    if (vertices.empty() || tris.empty()) { g_obj_err = std::string(path) + ": no geometry"; return ZDR_ERR_ARG; }
    const size_t nv = vertices.size(), nt = tris.size() / 3;
    float* verts = (float*)malloc(sizeof(float) * 8 * nv);
    int32_t* tri = (int32_t*)malloc(sizeof(int32_t) * 3 * nt);
    if (!verts || !tri) { free(verts); free(tri); g_obj_err = "out of host memory"; return ZDR_ERR_ARG; }
    for (size_t i = 0; i < nv; ++i)
        for (int k = 0; k < 8; ++k) verts[8 * i + k] = (float)vertices[i].v[k];
    memcpy(tri, tris.data(), sizeof(int32_t) * 3 * nt);
    if (std::isnan(vertices[0].v[5])) {
        // recompute_normal.py:4-51 as zdr_b200/load_obj.py states it: float32, un-normalised face normal added to
        // corner 0 of every triangle, then corner 1, then corner 2 (the order of the three np.add.at calls)
        std::vector<float> fn(3 * nt), acc(3 * nv, 0.0f);
        for (size_t t = 0; t < nt; ++t) {
            const float* p0 = verts + 8 * (size_t)tri[3 * t], *p1 = verts + 8 * (size_t)tri[3 * t + 1], *p2 = verts + 8 * (size_t)tri[3 * t + 2];
            const float a0 = p1[0] - p0[0], a1 = p1[1] - p0[1], a2 = p1[2] - p0[2];
            const float b0 = p2[0] - p0[0], b1 = p2[1] - p0[1], b2 = p2[2] - p0[2];
            const float m0 = a1 * b2, m1 = a2 * b1, m2 = a2 * b0, m3 = a0 * b2, m4 = a0 * b1, m5 = a1 * b0;
            fn[3 * t] = m0 - m1; fn[3 * t + 1] = m2 - m3; fn[3 * t + 2] = m4 - m5;
        }
        for (int k = 0; k < 3; ++k)
            for (size_t t = 0; t < nt; ++t) {
                float* a = &acc[3 * (size_t)tri[3 * t + k]];
                a[0] += fn[3 * t]; a[1] += fn[3 * t + 1]; a[2] += fn[3 * t + 2];
            }
        for (size_t i = 0; i < nv; ++i) {
            const float x = acc[3 * i], y = acc[3 * i + 1], z = acc[3 * i + 2];
            const float xx = x * x, yy = y * y, zz = z * z;
            float sum = xx + yy;
            sum = sum + zz;
            const float len = sqrtf(sum);
            verts[8 * i + 5] = x / len; verts[8 * i + 6] = y / len; verts[8 * i + 7] = z / len;
        }
    }
    *verts_out = verts; *nv_out = (int)nv; *tris_out = tri; *nt_out = (int)nt;
    return ZDR_OK;
}
