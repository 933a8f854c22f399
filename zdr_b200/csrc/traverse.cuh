// traverse.cuh — BVH2 traversal with the watertight ray/triangle test.
//
// Replaces accel.trace_closest / accel.trace_any (OptiX inside LuisaCompute; call sites
// collocated.py:16, direct.py:22,44,67, prb.py:25,59).  B200 has no RT cores, so this is a
// per-thread stack traversal over 64-byte nodes (two child boxes per node, fetched as four
// 16-byte vector loads) with the intersection test of Woop, Benthin, Wald, "Watertight
// Ray/Triangle Intersection", JCGT 2013.  The arithmetic of the triangle test is part of the
// numeric contract: the oracle evaluates the same expression per (ray, triangle), so the
// closest hit (smallest t, ties to the lowest scene-order triangle id) is identical no matter
// in which order the BVH visits the triangles.  The box test only has to be conservative.
#pragma once
#include "scene.cuh"

namespace zdr {

// node record: n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y), n1 = same for child 1,
//              n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z), n3 = (link0, link1, -, -) as int bits.
// link >= 0: inner node index; link < 0: leaf, ~link = (first_slot << 2) | (count - 1).
constexpr int kStackSize = 64;
constexpr int kLinkSentinel = 0x7fffffff;

struct RayPre {
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

ZDR_DEV float comp(float3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }

ZDR_DEV RayPre ray_pre(float3 d) {
    RayPre r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    r.kz = (ax >= ay && ax >= az) ? 0 : (ay >= az ? 1 : 2);
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    float dz = comp(d, r.kz);
    if (dz < 0.0f) { int t = r.kx; r.kx = r.ky; r.ky = t; }
    r.Sx = comp(d, r.kx) / dz;
    r.Sy = comp(d, r.ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

// true when the ray's line meets the triangle at t > tmin (t,u,v written; the caller applies the
// upper bound); watertight on shared edges.
ZDR_DEV bool tri_test(const RayPre& rp, float3 o, float3 p0, float3 p1, float3 p2, float tmin,
                      float& t_out, float& u_out, float& v_out) {
    float3 A = p0 - o, B = p1 - o, C = p2 - o;
    float Akz = comp(A, rp.kz), Bkz = comp(B, rp.kz), Ckz = comp(C, rp.kz);
    float Ax = fmaf(-rp.Sx, Akz, comp(A, rp.kx)), Ay = fmaf(-rp.Sy, Akz, comp(A, rp.ky));
    float Bx = fmaf(-rp.Sx, Bkz, comp(B, rp.kx)), By = fmaf(-rp.Sy, Bkz, comp(B, rp.ky));
    float Cx = fmaf(-rp.Sx, Ckz, comp(C, rp.kx)), Cy = fmaf(-rp.Sy, Ckz, comp(C, rp.ky));
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = U + V + W;
    if (det == 0.0f) return false;
    float Az = rp.Sz * Akz, Bz = rp.Sz * Bkz, Cz = rp.Sz * Ckz;
    float T = fmaf(W, Cz, fmaf(V, Bz, U * Az));
    float rcp = 1.0f / det;
    float t = T * rcp;
    if (!(t > tmin)) return false;
    t_out = t; u_out = V * rcp; v_out = W * rcp;
    return true;
}

struct TraceStats { unsigned nodes, tris; };

// conservative slab test; returns entry distance in tn
ZDR_DEV bool box_test(float lox, float hix, float loy, float hiy, float loz, float hiz, float3 o, float3 id,
                      float tmin, float tmax, float& tn) {
    float x0 = (lox - o.x) * id.x, x1 = (hix - o.x) * id.x;
    float y0 = (loy - o.y) * id.y, y1 = (hiy - o.y) * id.y;
    float z0 = (loz - o.z) * id.z, z1 = (hiz - o.z) * id.z;
    float tnear = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tmin));
    float tfar = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
    tn = tnear;
    return tnear * 0.9999995f <= tfar * 1.0000005f;
}

// Closest hit (ANY=false) or occlusion query (ANY=true; returns hit.inst >= 0 when occluded).
template <bool ANY, bool STATS>
ZDR_DEV Hit trace(const SceneView& sc, const Ray& ray, TraceStats& st) {
    Hit hit;
    hit.inst = -1; hit.prim = -1; hit.slot = -1; hit.u = 0.0f; hit.v = 0.0f; hit.t = 0.0f;
    const RayPre rp = ray_pre(ray.d);
    const float3 id = f3(1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z);
    float best = ray.tmax;
    int best_gid = 0x7fffffff;
    int stack[kStackSize];
    int sp = 0;
    stack[sp++] = kLinkSentinel;
    int link = sc.root_link;
    while (link != kLinkSentinel) {
        if (link >= 0) {
            const float4* n = sc.nodes + 4 * (size_t)link;
            const float4 n0 = __ldg(n), n1 = __ldg(n + 1), n2 = __ldg(n + 2), n3 = __ldg(n + 3);
            if (STATS) st.nodes++;
            float t0, t1;
            // ties on t must stay reachable: cull with <= best
            const bool h0 = box_test(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, ray.o, id, ray.tmin, best, t0);
            const bool h1 = box_test(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, ray.o, id, ray.tmin, best, t1);
            const int l0 = __float_as_int(n3.x), l1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = t1 < t0;
                stack[sp++] = swap ? l0 : l1;
                link = swap ? l1 : l0;
            } else if (h0) {
                link = l0;
            } else if (h1) {
                link = l1;
            } else {
                link = stack[--sp];
            }
        } else {
            const int code = ~link;
            const int first = code >> 2, count = (code & 3) + 1;
            for (int k = 0; k < count; ++k) {
                const float4* tp = sc.tris + 3 * (size_t)(first + k);
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (STATS) st.tris++;
                float t, u, v;
                if (tri_test(rp, ray.o, f3(a.x, a.y, a.z), f3(b.x, b.y, b.z), f3(c.x, c.y, c.z), ray.tmin, t, u, v)) {
                    if (ANY) {
                        if (t < ray.tmax) { hit.inst = 0; return hit; }
                        continue;
                    }
                    const int gid = __float_as_int(a.w);
                    // smallest t wins; equal t goes to the lowest scene-order triangle id, which is
                    // what the oracle's in-order scan with a strict '<' keeps
                    if (t < best || (hit.inst >= 0 && t == best && gid < best_gid)) {
                        best = t; best_gid = gid;
                        hit.t = t; hit.u = u; hit.v = v; hit.slot = first + k;
                        hit.inst = __float_as_int(b.w); hit.prim = __float_as_int(c.w);
                    }
                }
            }
            link = stack[--sp];
        }
    }
    return hit;
}

}  // namespace zdr
