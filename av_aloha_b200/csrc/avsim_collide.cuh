// avsim_collide.cuh -- fp32 narrowphase for the warp-per-environment step kernel.
//
// Replaces MuJoCo's mj_collision narrowphase [third-party; reached through Physics.step, reference env.py:218].
// Primitive pairs (sphere-sphere, sphere-box, box-box) run one candidate pair per lane; every other convex pair
// (mesh hulls, cylinder) runs Minkowski Portal Refinement cooperatively on the whole warp: control flow is
// warp-uniform and the hull support map is an exhaustive arg-max strided over the 32 lanes (hull vertices are
// float4 in L2-resident model memory, so one support query is ceil(nvert/32) coalesced 512-byte reads).
// Contact convention: normal from geom1 to geom2, dist < 0 is penetration, pos is the mid-surface point.
#pragma once
#include "avsim_dev.h"
#include "avsim_math.cuh"

#define AV_MPR_TOL 1e-6f
#define AV_MPR_ITERS 50
#ifndef AV_SUP_REDUX
#define AV_SUP_REDUX 0   // 1: REDUX argmax + register-carried winner in the hull support scan instead of the shuffle butterfly + reload.
                         // Same results (emulator parity green); measured SLOWER on B200: 40.3 vs 39.95 ms/step (profiles/r1_summary.md)
#endif
#define AV_SUP_UNROLL 11   // 11 x 32 = 352 vertices per chunk: the finger hulls (345 / 347 vertices) scan in one chunk

struct Shape {
    int type, nvert;
    V3 pos, size;
    M3 mat;
    const float4 *vert;
};

struct PrimOut {   // <= 8 points sharing one normal
    int n;
    V3 nrm;
    float dist[8];
    V3 pos[8];
};

__device__ __forceinline__ void make_frame(V3 n, V3 &t1, V3 &t2) {
    V3 e = fabsf(n.y) < 0.5f ? v3(0, 1, 0) : v3(0, 0, 1);
    t1 = normalized(e - n * dot(e, n));
    t2 = cross(n, t1);
}

// ------------------------------------------------------------------ closed forms (one pair per lane)
__device__ inline void collide_sphere_sphere(const Shape &A, const Shape &B, PrimOut &o) {
    V3 d = B.pos - A.pos;
    float len = norm(d), dist = len - A.size.x - B.size.x;
    o.n = 0;
    if (dist >= 0) return;
    d = len < AV_MINVAL ? v3(0, 0, 1) : d * (1.0f / len);
    o.n = 1; o.nrm = d; o.dist[0] = dist;
    o.pos[0] = A.pos + d * (A.size.x + 0.5f * dist);
}

// A sphere, B box; normal from the sphere to the box
__device__ inline void collide_sphere_box(const Shape &A, const Shape &B, PrimOut &o) {
    V3 c = mulT(B.mat, A.pos - B.pos);
    V3 q = v3(fminf(fmaxf(c.x, -B.size.x), B.size.x), fminf(fmaxf(c.y, -B.size.y), B.size.y),
              fminf(fmaxf(c.z, -B.size.z), B.size.z));
    bool inside = (q.x == c.x) && (q.y == c.y) && (q.z == c.z);
    float r = A.size.x, dist;
    V3 nl;
    o.n = 0;
    if (!inside) {
        V3 dv = c - q;
        float len = norm(dv);
        dist = len - r;
        if (dist >= 0) return;
        nl = dv * (-1.0f / len);
    } else {
        float dx = B.size.x - fabsf(c.x), dy = B.size.y - fabsf(c.y), dz = B.size.z - fabsf(c.z);
        nl = v3(0, 0, 0);
        float bd;
        if (dx <= dy && dx <= dz) { bd = dx; nl.x = c.x >= 0 ? -1.f : 1.f; q.x = c.x >= 0 ? B.size.x : -B.size.x; }
        else if (dy <= dz) { bd = dy; nl.y = c.y >= 0 ? -1.f : 1.f; q.y = c.y >= 0 ? B.size.y : -B.size.y; }
        else { bd = dz; nl.z = c.z >= 0 ? -1.f : 1.f; q.z = c.z >= 0 ? B.size.z : -B.size.z; }
        dist = -bd - r;
    }
    V3 nw = mul(B.mat, nl), qw = mul(B.mat, q) + B.pos;
    o.n = 1; o.nrm = nw; o.dist[0] = dist;
    o.pos[0] = qw - nw * (0.5f * dist);   // midpoint of (box surface point, deepest sphere point): |dist| / 2 beyond the surface along nw
}

// ------------------------------------------------------------------ box-box: 15-axis SAT + face clipping / edge-edge
__device__ inline int clip_poly(float (*poly)[3], int n, int axis, float lim) {
    for (int pass = 0; pass < 2; pass++) {
        float sgn = pass ? -1.f : 1.f, outp[16][3];
        int no = 0;
        for (int i = 0; i < n; i++) {
            const float *a = poly[i], *b = poly[(i + 1) % n];
            float da = sgn * a[axis] - lim, db = sgn * b[axis] - lim;
            if (da <= 0) { outp[no][0] = a[0]; outp[no][1] = a[1]; outp[no][2] = a[2]; no++; }
            if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
                float t = da / (da - db);
                for (int k = 0; k < 3; k++) outp[no][k] = a[k] + t * (b[k] - a[k]);
                no++;
            }
        }
        n = no;
        for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) poly[i][k] = outp[i][k];
        if (!n) break;
    }
    return n;
}

__device__ inline void collide_box_box(const Shape &A, const Shape &B, PrimOut &o) {
    V3 a[3], b[3];
    float R[3][3], Q[3][3], dA[3], dB[3];
    float hA[3] = {A.size.x, A.size.y, A.size.z}, hB[3] = {B.size.x, B.size.y, B.size.z};
    for (int k = 0; k < 3; k++) { a[k] = colm(A.mat, k); b[k] = colm(B.mat, k); }
    V3 d = B.pos - A.pos;
    for (int i = 0; i < 3; i++) {
        dA[i] = dot(d, a[i]); dB[i] = dot(d, b[i]);
        for (int j = 0; j < 3; j++) { R[i][j] = dot(a[i], b[j]); Q[i][j] = fabsf(R[i][j]); }
    }
    o.n = 0;
    float best = -1e30f;
    int code = -1;
    for (int i = 0; i < 3; i++) {
        float s = fabsf(dA[i]) - (hA[i] + hB[0] * Q[i][0] + hB[1] * Q[i][1] + hB[2] * Q[i][2]);
        if (s > 0) return;
        if (s > best) { best = s; code = i; }
    }
    for (int j = 0; j < 3; j++) {
        float s = fabsf(dB[j]) - (hB[j] + hA[0] * Q[0][j] + hA[1] * Q[1][j] + hA[2] * Q[2][j]);
        if (s > 0) return;
        if (s > best) { best = s; code = 3 + j; }
    }
    float ebest = -1e30f;
    int ecode = -1;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float len2 = 1.f - R[i][j] * R[i][j];
            if (len2 < 1e-4f) continue;   // edges within 0.6 deg of parallel (same rule as the oracle): the cross axis is a 0/0 in fp32
            int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
            float il = rsqrtf(len2);
            float s = (fabsf(dA[i2] * R[i1][j] - dA[i1] * R[i2][j]) -
                       (hA[i1] * Q[i2][j] + hA[i2] * Q[i1][j] + hB[j1] * Q[i][j2] + hB[j2] * Q[i][j1])) * il;
            if (s > 0) return;
            if (s > ebest) { ebest = s; ecode = 3 * i + j; }
        }
    if (ecode >= 0 && ebest > 0.95f * best + 1e-9f) {
        int i = ecode / 3, j = ecode % 3;
        V3 L = normalized(cross(a[i], b[j]));
        if (dot(L, d) < 0) L = -L;
        V3 pa = A.pos, pb = B.pos;
        for (int k = 0; k < 3; k++) {
            if (k != i) pa = pa + a[k] * ((dot(L, a[k]) > 0 ? 1.f : -1.f) * hA[k]);
            if (k != j) pb = pb + b[k] * ((dot(L, b[k]) > 0 ? -1.f : 1.f) * hB[k]);
        }
        V3 w = pa - pb;
        float ab = R[i][j], wa = dot(w, a[i]), wb = dot(w, b[j]), den = 1.f - ab * ab;
        float s = (ab * wb - wa) / den, t = (wb - ab * wa) / den;
        s = fminf(fmaxf(s, -hA[i]), hA[i]); t = fminf(fmaxf(t, -hB[j]), hB[j]);   // the closest points lie ON the two edges
        o.n = 1; o.nrm = L; o.dist[0] = ebest;
        o.pos[0] = ((pa + a[i] * s) + (pb + b[j] * t)) * 0.5f;
        return;
    }
    bool refA = code < 3;
    const Shape &Rf = refA ? A : B;
    const Shape &In = refA ? B : A;
    const V3 *ra = refA ? a : b, *ia = refA ? b : a;
    const float *hR = refA ? hA : hB, *hI = refA ? hB : hA;
    int ri = code % 3;
    V3 dref = In.pos - Rf.pos;
    float sgn = dot(dref, ra[ri]) >= 0 ? 1.f : -1.f;
    V3 nref = ra[ri] * sgn;
    int ij = 0;
    float bestdot = -1.f;
    for (int k = 0; k < 3; k++) {
        float dt = fabsf(dot(nref, ia[k]));
        if (dt > bestdot) { bestdot = dt; ij = k; }
    }
    float isgn = dot(nref, ia[ij]) > 0 ? -1.f : 1.f;
    int iu = (ij + 1) % 3, iv = (ij + 2) % 3, ru = (ri + 1) % 3, rv = (ri + 2) % 3;
    float poly[16][3];
    for (int c = 0; c < 4; c++) {
        float su = (c == 0 || c == 3) ? 1.f : -1.f, sv = (c < 2) ? 1.f : -1.f;
        V3 p = dref + ia[ij] * (isgn * hI[ij]) + ia[iu] * (su * hI[iu]) + ia[iv] * (sv * hI[iv]);
        poly[c][0] = dot(p, ra[ru]); poly[c][1] = dot(p, ra[rv]); poly[c][2] = dot(p, nref);
    }
    int n = clip_poly(poly, 4, 0, hR[ru]);
    if (n) n = clip_poly(poly, n, 1, hR[rv]);
    int nc = 0;
    for (int c = 0; c < n && nc < 8; c++) {
        float sep = poly[c][2] - hR[ri];
        if (sep > 0) continue;
        o.dist[nc] = sep;
        o.pos[nc] = Rf.pos + ra[ru] * poly[c][0] + ra[rv] * poly[c][1] + nref * (poly[c][2] - 0.5f * sep);
        nc++;
    }
    o.n = nc;
    o.nrm = refA ? nref : -nref;
}

// ------------------------------------------------------------------ warp-cooperative MPR
// The support points are fp32 (vertex scan, one per lane-strided chunk); everything that DECIDES -- which side of a portal
// edge the origin ray passes, when the portal has converged, the distance of the origin to the final portal -- runs in fp64 on
// the exact difference of the two fp32 support points.  In fp32 those triple products (three 5 cm vectors that differ by
// millimetres) keep 2-4 significant digits, and whenever the origin ray leaves the Minkowski difference within ~1e-3 of an
// edge the portal walked onto the neighbouring face: 5 % of the hull contacts of the bench workload came out with another
// normal (up to 60 degrees off) and up to 0.9 mm another depth than the fp64 oracle (profiles/r2_newton_parity.txt, the MPR lines).  The decisions
// are a few dozen scalar operations per iteration next to a scan over hundreds of vertices: B200's fp64 pipe makes them free.
struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) { D3 r = {x, y, z}; return r; }
__device__ __forceinline__ D3 d3(V3 a) { return d3((double)a.x, (double)a.y, (double)a.z); }
__device__ __forceinline__ V3 f3(D3 a) { return v3((float)a.x, (float)a.y, (float)a.z); }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 operator-(D3 a) { return d3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ D3 operator*(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) { return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double norm(D3 a) { return sqrt(dot(a, a)); }
// unit vector: fp32 reciprocal square root + two Newton steps in fp64 (no fp64 sqrt / division on the portal loop's critical path)
__device__ __forceinline__ D3 normalized(D3 a) {
    double n2 = dot(a, a);
    if (!(n2 > 1e-60)) return a;
    double r = (double)rsqrtf((float)n2);
    r = r * (1.5 - 0.5 * n2 * r * r);
    r = r * (1.5 - 0.5 * n2 * r * r);
    return a * r;
}
struct Sup {
    D3 v;        // v1 - v2, exact
    V3 v1, v2;   // support points on the two shapes
};

// support point of S in world direction dir; uniform across the warp
__device__ inline V3 support_world(const Shape &S, V3 dir, int lane) {
    V3 dl = mulT(S.mat, dir), pl;
    if (S.type == AV_GEOM_MESH) {
        float bd = -3.0e38f;
        int bi = 0x7fffffff;
#if AV_SUP_REDUX
        float bx = 0.f, by = 0.f, bz = 0.f;     // the lane's best vertex travels with its score: no dependent reload of the winner
#endif
        // up to AV_SUP_UNROLL independent 512-byte row loads in flight per lane before the first use: the scan is bound by L2 latency
        // (ncu: one third of all stall samples sat on the dependent load->FMA of the rolled loop), not by bandwidth
        for (int base = lane; base < S.nvert; base += 32 * AV_SUP_UNROLL) {
            float4 pv[AV_SUP_UNROLL];
#pragma unroll
            for (int u = 0; u < AV_SUP_UNROLL; u++) {
                int i = base + 32 * u;
                pv[u] = ldg4(S.vert + (i < S.nvert ? i : lane));     // unconditional: a predicated load breaks the batching
            }
#pragma unroll
            for (int u = 0; u < AV_SUP_UNROLL; u++) {
                int i = base + 32 * u;
                float dt = pv[u].x * dl.x + pv[u].y * dl.y + pv[u].z * dl.z;
#if AV_SUP_REDUX
                if (i < S.nvert && dt > bd) { bd = dt; bi = i; bx = pv[u].x; by = pv[u].y; bz = pv[u].z; }
#else
                if (i < S.nvert && dt > bd) { bd = dt; bi = i; }
#endif
            }
        }
#if AV_SUP_REDUX
        // argmax by two REDUX ops on an order-preserving key (ties -> smallest index, as warp_argmax), then the owner lane
        // (vertex i lives in lane i & 31) broadcasts the coordinates it already holds
        bd += 0.0f;                                                  // -0.0 -> +0.0: the key order must match float compare
        warp_argmax_redux(bd, bi);
        const int owner = bi & 31;
        pl = v3(__shfl_sync(AV_FULL, bx, owner), __shfl_sync(AV_FULL, by, owner), __shfl_sync(AV_FULL, bz, owner));
#else
        warp_argmax(bd, bi);
        float4 p = ldg4(S.vert + bi);
        pl = v3(p.x, p.y, p.z);
#endif
    } else if (S.type == AV_GEOM_BOX) {
        pl = v3(dl.x >= 0 ? S.size.x : -S.size.x, dl.y >= 0 ? S.size.y : -S.size.y, dl.z >= 0 ? S.size.z : -S.size.z);
    } else if (S.type == AV_GEOM_SPHERE) {
        float n = norm(dl);
        pl = n < AV_MINVAL ? v3(0, 0, 0) : dl * (S.size.x / n);
    } else {  // cylinder: radius size.x, half height size.y along local z
        float n = sqrtf(dl.x * dl.x + dl.y * dl.y);
        pl = n > AV_MINVAL ? v3(dl.x / n * S.size.x, dl.y / n * S.size.x, 0) : v3(0, 0, 0);
        pl.z = dl.z >= 0 ? S.size.y : -S.size.y;
    }
    return mul(S.mat, pl) + S.pos;
}
// Deliberately NOT inlined (and mpr_penetration has one call site): the step kernel's instruction footprint is what
// bounds it (ncu: 'no_instruction' stalls), and MPR inlined at every call site was 190 KB of SASS.
__device__ __noinline__ void mpr_support_ni(const Shape &A, const Shape &B, float dx, float dy, float dz, int lane, Sup &s) {
    V3 dir = v3(dx, dy, dz);
    s.v1 = support_world(A, dir, lane);
    s.v2 = support_world(B, -dir, lane);
    s.v = d3(s.v1) - d3(s.v2);
}
__device__ __forceinline__ Sup mpr_support(const Shape &A, const Shape &B, D3 dir, int lane) {
    Sup s;
    mpr_support_ni(A, B, (float)dir.x, (float)dir.y, (float)dir.z, lane, s);
    return s;
}
__device__ __forceinline__ D3 portal_dir(const Sup *p) { return normalized(cross(p[2].v - p[1].v, p[3].v - p[1].v)); }
__device__ __forceinline__ void expand_portal(Sup *p, const Sup &v4) {
    D3 v4v0 = cross(v4.v, p[0].v);
    if (dot(p[1].v, v4v0) > 0) {
        if (dot(p[2].v, v4v0) > 0) p[1] = v4; else p[3] = v4;
    } else {
        if (dot(p[3].v, v4v0) > 0) p[2] = v4; else p[1] = v4;
    }
}
__device__ __forceinline__ bool reach_tolerance(const Sup *p, const Sup &v4, D3 dir) {
    double dv4 = dot(v4.v, dir);
    double dt = fmin(fmin(dv4 - dot(p[1].v, dir), dv4 - dot(p[2].v, dir)), dv4 - dot(p[3].v, dir));
    return dt <= (double)AV_MPR_TOL;
}
__device__ inline double origin_tri_dist2(D3 a, D3 b, D3 c, D3 &wit) {
    D3 ab = b - a, ac = c - a, ap = -a;
    double d1 = dot(ab, ap), d2 = dot(ac, ap), s, t;
    if (d1 <= 0 && d2 <= 0) { s = 0; t = 0; }
    else {
        double d3_ = dot(ab, -b), d4 = dot(ac, -b);
        if (d3_ >= 0 && d4 <= d3_) { s = 1; t = 0; }
        else {
            double vc = d1 * d4 - d3_ * d2;
            if (vc <= 0 && d1 >= 0 && d3_ <= 0) { s = d1 / (d1 - d3_); t = 0; }
            else {
                double d5 = dot(ab, -c), d6 = dot(ac, -c);
                if (d6 >= 0 && d5 <= d6) { s = 0; t = 1; }
                else {
                    double vb = d5 * d2 - d1 * d6;
                    if (vb <= 0 && d2 >= 0 && d6 <= 0) { s = 0; t = d2 / (d2 - d6); }
                    else {
                        double va = d3_ * d6 - d5 * d4;
                        if (va <= 0 && (d4 - d3_) >= 0 && (d5 - d6) >= 0) { t = (d4 - d3_) / ((d4 - d3_) + (d5 - d6)); s = 1 - t; }
                        else { double den = 1.0 / (va + vb + vc); s = vb * den; t = vc * den; }
                    }
                }
            }
        }
    }
    wit = a + ab * s + ac * t;
    return dot(wit, wit);
}
__device__ inline V3 find_pos(const Sup *p) {
    D3 dir = portal_dir(p);
    double b[4];
    b[0] = dot(cross(p[1].v, p[2].v), p[3].v);
    b[1] = dot(cross(p[3].v, p[2].v), p[0].v);
    b[2] = dot(cross(p[0].v, p[1].v), p[3].v);
    b[3] = dot(cross(p[2].v, p[1].v), p[0].v);
    double sum = b[0] + b[1] + b[2] + b[3];
    if (sum <= 0) {
        b[0] = 0;
        b[1] = dot(cross(p[2].v, p[3].v), dir);
        b[2] = dot(cross(p[3].v, p[1].v), dir);
        b[3] = dot(cross(p[1].v, p[2].v), dir);
        sum = b[1] + b[2] + b[3];
    }
    double inv = 1.0 / sum;
    D3 p1 = d3(0, 0, 0), p2 = d3(0, 0, 0);
    for (int i = 0; i < 4; i++) { p1 = p1 + d3(p[i].v1) * b[i]; p2 = p2 + d3(p[i].v2) * b[i]; }
    return f3((p1 + p2) * (0.5 * inv));
}

// returns true (warp-uniform) and depth / direction A->B / position when the shapes intersect
__device__ __noinline__ bool mpr_penetration(const Shape &A, const Shape &B, int lane, float &depth, V3 &dir_out, V3 &pos) {
    Sup p[4], v4;
    D3 dir;
    p[0].v1 = A.pos; p[0].v2 = B.pos; p[0].v = d3(A.pos) - d3(B.pos);
    if (norm(p[0].v) < 1e-9) p[0].v.x += 1e-6;
    dir = normalized(-p[0].v);
    p[1] = mpr_support(A, B, dir, lane);
    if (dot(p[1].v, dir) <= 0) return false;
    dir = cross(p[0].v, p[1].v);
    if (norm(dir) < 1e-12) {
        pos = (p[1].v1 + p[1].v2) * 0.5f;
        if (norm(p[1].v) < 1e-9) { depth = 0; dir_out = v3(0, 0, 0); return true; }
        depth = (float)norm(p[1].v);
        dir_out = f3(normalized(p[1].v));
        return true;
    }
    dir = normalized(dir);
    p[2] = mpr_support(A, B, dir, lane);
    if (dot(p[2].v, dir) <= 0) return false;
    dir = normalized(cross(p[1].v - p[0].v, p[2].v - p[0].v));
    if (dot(dir, p[0].v) > 0) { Sup t = p[1]; p[1] = p[2]; p[2] = t; dir = -dir; }
    for (int it = 0;; it++) {
        if (it > 100) return false;
        p[3] = mpr_support(A, B, dir, lane);
        if (dot(p[3].v, dir) <= 0) return false;
        bool cont = false;
        if (dot(cross(p[1].v, p[3].v), p[0].v) < 0) { p[2] = p[3]; cont = true; }
        if (!cont && dot(cross(p[3].v, p[2].v), p[0].v) < 0) { p[1] = p[3]; cont = true; }
        if (!cont) break;
        dir = normalized(cross(p[1].v - p[0].v, p[2].v - p[0].v));
    }
    for (int it = 0;; it++) {
        dir = portal_dir(p);
        if (dot(dir, p[1].v) >= 0) break;
        v4 = mpr_support(A, B, dir, lane);
        if (dot(v4.v, dir) < 0 || reach_tolerance(p, v4, dir) || it > 100) return false;
        expand_portal(p, v4);
    }
    for (int it = 0;; it++) {
        dir = portal_dir(p);
        v4 = mpr_support(A, B, dir, lane);
        if (reach_tolerance(p, v4, dir) || it > AV_MPR_ITERS) {
            D3 wit;
            double d2 = origin_tri_dist2(p[1].v, p[2].v, p[3].v, wit);
            depth = (float)sqrt(d2);
            dir_out = f3(depth >= 1e-9f ? normalized(wit) : dir);
            pos = find_pos(p);
            return true;
        }
        expand_portal(p, v4);
    }
}

__device__ inline M3 rot_axis_angle(V3 ax, float ang) {
    float c = cosf(ang), s = sinf(ang), t = 1 - c, x = ax.x, y = ax.y, z = ax.z;
    M3 R;
    R.m[0] = t * x * x + c; R.m[1] = t * x * y - s * z; R.m[2] = t * x * z + s * y;
    R.m[3] = t * x * y + s * z; R.m[4] = t * y * y + c; R.m[5] = t * y * z - s * x;
    R.m[6] = t * x * z - s * y; R.m[7] = t * y * z + s * x; R.m[8] = t * z * z + c;
    return R;
}

// warp-cooperative convex pair: up to 5 points (multiccd) sharing the unperturbed normal.  q = -1 is the unperturbed
// run, q = 0..3 the +-1e-3 rad rotations about the two tangent axes through the first contact point.
__device__ inline void collide_convex(const Shape &A, const Shape &B, bool multiccd, int lane, PrimOut &o) {
    o.n = 0;
    V3 t1 = v3(0, 0, 0), t2 = v3(0, 0, 0), pos0 = v3(0, 0, 0);
    Shape A2 = A, B2 = B;
    int nq = multiccd ? 4 : 0;
    for (int q = -1; q < nq; q++) {
        if (q >= 0) {
            V3 ax = q < 2 ? t1 : t2;
            float ang = (q & 1) ? -1e-3f : 1e-3f;
            M3 Rp = rot_axis_angle(ax, ang), Rm = rot_axis_angle(ax, -ang);
            A2.mat = mul(Rp, A.mat); A2.pos = pos0 + mul(Rp, A.pos - pos0);
            B2.mat = mul(Rm, B.mat); B2.pos = pos0 + mul(Rm, B.pos - pos0);
        }
        float dp;
        V3 dr, ps;
        bool hit = mpr_penetration(A2, B2, lane, dp, dr, ps) && norm(dr) >= 0.5f;
        if (q < 0) {
            if (!hit) return;
            o.n = 1; o.nrm = dr; o.dist[0] = -dp; o.pos[0] = ps;
            pos0 = ps;
            make_frame(dr, t1, t2);
            continue;
        }
        if (!hit) continue;
        bool dup = false;
        for (int c = 0; c < o.n; c++) dup = dup || norm(ps - o.pos[c]) < 1e-4f;
        if (dup) continue;
        o.dist[o.n] = -dp; o.pos[o.n] = ps; o.n++;
    }
}

// One MPR run of a convex pair, the unit of work the lockstep kernel pools across a block's environments: q = -1 the
// unperturbed shapes; q = 0..3 (multiccd) the shapes rotated by +-1e-3 rad about the tangent axes of the unperturbed
// normal nrm0 through the unperturbed contact point pos0.  Same arithmetic as collide_convex above, run by run.
__device__ inline bool collide_convex_run(const Shape &A, const Shape &B, int q, V3 pos0, V3 nrm0, int lane, float &dist, V3 &pos, V3 &nrm) {
    Shape A2 = A, B2 = B;
    if (q >= 0) {
        V3 t1, t2;
        make_frame(nrm0, t1, t2);
        V3 ax = q < 2 ? t1 : t2;
        float ang = (q & 1) ? -1e-3f : 1e-3f;
        M3 Rp = rot_axis_angle(ax, ang), Rm = rot_axis_angle(ax, -ang);
        A2.mat = mul(Rp, A.mat); A2.pos = pos0 + mul(Rp, A.pos - pos0);
        B2.mat = mul(Rm, B.mat); B2.pos = pos0 + mul(Rm, B.pos - pos0);
    }
    float dp;
    bool hit = mpr_penetration(A2, B2, lane, dp, nrm, pos) && norm(nrm) >= 0.5f;
    dist = -dp;
    return hit;
}

// conservative oriented-box test on the 6 face axes of the two local bounding boxes
__device__ inline bool obb_separated(V3 hA, const Shape &A, V3 hB, const Shape &B) {
    V3 d = B.pos - A.pos;
    for (int i = 0; i < 3; i++) {
        V3 ai = colm(A.mat, i);
        float rb = hB.x * fabsf(dot(ai, colm(B.mat, 0))) + hB.y * fabsf(dot(ai, colm(B.mat, 1))) + hB.z * fabsf(dot(ai, colm(B.mat, 2)));
        if (fabsf(dot(d, ai)) > comp(hA, i) + rb) return true;
    }
    for (int j = 0; j < 3; j++) {
        V3 bj = colm(B.mat, j);
        float ra = hA.x * fabsf(dot(bj, colm(A.mat, 0))) + hA.y * fabsf(dot(bj, colm(A.mat, 1))) + hA.z * fabsf(dot(bj, colm(A.mat, 2)));
        if (fabsf(dot(d, bj)) > comp(hB, j) + ra) return true;
    }
    return false;
}
