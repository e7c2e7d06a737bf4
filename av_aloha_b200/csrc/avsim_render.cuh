// avsim_render.cuh -- K9: multi-camera renderer, writes uint8 [B][ncam][H][W][3] (what physics.render(h, w, camera_id)
// returns per camera in get_obs, reference gym_guided_vision/env.py:180-188,195-200).
//
// Fidelity (stated, not hidden): this is a ray caster over the PHYSICS geoms, not a rasteriser of the 588 k visual
// triangles.  Task objects, table and finger pads are drawn exactly as their primitives; every mesh geom (robot links, frame
// extrusions) is drawn as the 26-DOP of its convex hull -- the intersection of 13 slabs (3 axes, 6 face diagonals, 4 body
// diagonals) whose extents the model compiler takes from the hull vertices -- in the colour of the visual mesh it stands for:
// a bevelled convex stand-in whose silhouette covers the hull's to within a few per cent (tests/test_gpu_render.py measures the
// IoU against the projected hulls; round 1 drew the 6-faced bounding box).  Lighting as the scene defines it: headlight ambient
// 0.3 + diffuse 0.6 (scene.xml:9) and the directional light pointing down (scene.xml:50, diffuse 0.7), no shadows, no specular;
// the table top carries its diffuse texture (scene.xml:30-32, box-filtered to 128 x 128); gradient sky.  The reference's own
// renders are declared non-deterministic (gym_guided_vision/__init__.py:92-94) and no GL context exists here, so there is no
// pixel oracle: tests check camera geometry, silhouettes, determinism and coverage, not image equality.
//
// Two kernels:
//   avsim_render_prep_kernel  warp per environment: forward kinematics from qpos, world pose of every geom and camera
//                             -> rpose[B][ngeom + ncam_all][12] (pos 3 | rotation 9), and the screen rectangle of every
//                             geom's oriented box in every requested camera -> rrect[B][ncam][ngeom]; L2 resident
//   avsim_render_kernel       block per (64 x 32 pixel region, camera, environment): culls geoms by their screen-space
//                             8-DOP into shared memory, ordered front to back; each warp then shades 8 x 4 pixel blocks: a
//                             ballot culls the region's candidates against the block, one primary ray per lane over the
//                             survivors with a depth early-out, shade, stage the region in shared memory and write it as
//                             32-bit words (192 contiguous bytes per region row).  HBM-write bound in principle
//                             (H*W*3 bytes per image), issue-bound in practice (profiles/r2_render.txt).
#pragma once
#include "avsim_step.cuh"

#define AV_RT_W 64         // a block renders a 64 x 32 pixel region ...
#define AV_RT_H 32
#define AV_RT_THREADS 256   // ... with 8 warps, each shading 8 x 4 pixel blocks
#define AV_RT_MAXG 64      // candidate geoms per region (two ballot words)

__global__ void __launch_bounds__(32) avsim_render_prep_kernel(const __grid_constant__ DevModel m, const __grid_constant__ BatchState B,
                                                               float *__restrict__ rpose, float4 *__restrict__ rrect, int ncam_all,
                                                               const int *__restrict__ cam_body, const float *__restrict__ cam_pos,
                                                               const float *__restrict__ cam_quat, const float *__restrict__ cam_fovy,
                                                               const int *__restrict__ cam_ids, int ncam, int H, int W) {
    EnvS &S = *reinterpret_cast<EnvS *>(av_smem_raw);
    int lane = threadIdx.x;
    for (int env = blockIdx.x; env < B.num_envs; env += gridDim.x) {
        env_load(m, B, S, env, lane);
        stage_kinematics(m, S, lane);
        float *out = rpose + (size_t)env * (m.ngeom + ncam_all) * 12;
        for (int g = lane; g < m.ngeom; g += 32) {
            V3 p;
            M3 R;
            if (m.geom_static[g]) { p = ld3(m.geom_xpos0 + 3 * g); R = ldm3(m.geom_xmat0 + 9 * g); }
            else {
                int b = m.geom_body[g];
                M3 Rb = body_mat(S, b);
                p = ld3(S.xpos + 3 * b) + mul(Rb, ld3(m.geom_pos + 3 * g));
                R = mul(Rb, ldm3(m.geom_mat + 9 * g));
            }
            st3(out + 12 * g, p);
            stm3(out + 12 * g + 3, R);
        }
        for (int c = lane; c < ncam_all; c += 32) {
            int b = cam_body[c];
            M3 Rb = body_mat(S, b);
            V3 p = ld3(S.xpos + 3 * b) + mul(Rb, ld3(cam_pos + 3 * c));
            M3 R = mul(Rb, q2m(qnormalize(ldq(cam_quat + 4 * c))));
            st3(out + 12 * (m.ngeom + c), p);
            stm3(out + 12 * (m.ngeom + c) + 3, R);
        }
        __syncwarp();
        // screen-space bounding rectangle of every geom's oriented box in every requested camera: the tile cull of the
        // render kernel is then four compares per geom (bounding spheres let the long frame extrusions and the table
        // into every tile).  Empty = (1, 1, 0, 0); a box that straddles the camera plane gets the whole image.
        for (int ci = 0; ci < ncam; ci++) {
            int cam = cam_ids[ci];
            const float *cp = out + 12 * (m.ngeom + cam);
            V3 o = ld3(cp);
            M3 Rc = ldm3(cp + 3);
            float th = tanf(0.5f * cam_fovy[cam] * 0.017453292519943295f), aspect = (float)W / (float)H;
            for (int g = lane; g < m.ngeom; g += 32) {
                float4 rect = make_float4(1.f, 1.f, 0.f, 0.f), diag = make_float4(-1e30f, -1e30f, 1e30f, 1e30f);
                float zmin = 1e30f;
                if (m.geom_visible[g]) {
                    V3 c = ld3(out + 12 * g);
                    M3 R = ldm3(out + 12 * g + 3);
                    int ty = m.geom_type[g];
                    V3 h = ty == AV_GEOM_MESH ? ld3(m.geom_aabb + 3 * g) : ld3(m.geom_size + 3 * g);
                    if (ty == AV_GEOM_SPHERE) h = v3(h.x, h.x, h.x);
                    if (ty == AV_GEOM_CYLINDER) h = v3(h.x, h.x, h.y);
                    // corners in camera space; the part of the box in front of the near plane z = -zn is the convex hull of
                    // the in-front corners and the points where box edges cross that plane
                    const float zn = 0.02f, sx = 0.5f * W / (th * aspect), sy = 0.5f * H / th;
                    float x0 = 1e30f, y0 = 1e30f, x1 = -1e30f, y1 = -1e30f, u0 = 1e30f, u1 = -1e30f, v0 = 1e30f, v1 = -1e30f;
                    V3 pc[8];
                    int nfront = 0;
                    auto grow = [&](float x, float y) {
                        x0 = fminf(x0, x); x1 = fmaxf(x1, x); y0 = fminf(y0, y); y1 = fmaxf(y1, y);
                        u0 = fminf(u0, x + y); u1 = fmaxf(u1, x + y); v0 = fminf(v0, x - y); v1 = fmaxf(v1, x - y);
                    };
                    for (int k = 0; k < 8; k++) {
                        V3 pw = c + mul(R, v3((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z));
                        pc[k] = mulT(Rc, pw - o);
                        zmin = fminf(zmin, -pc[k].z);
                        if (pc[k].z <= -zn) {
                            nfront++;
                            float iz = -1.0f / pc[k].z;
                            grow(0.5f * W + pc[k].x * iz * sx, 0.5f * H - pc[k].y * iz * sy);
                        }
                    }
                    if (nfront > 0 && nfront < 8)
                        for (int k = 0; k < 8; k++)
                            for (int bit = 1; bit < 8; bit <<= 1) {
                                int j = k ^ bit;
                                if (j < k || (pc[k].z <= -zn) == (pc[j].z <= -zn)) continue;
                                float t = (-zn - pc[k].z) / (pc[j].z - pc[k].z);
                                float px_ = pc[k].x + t * (pc[j].x - pc[k].x), py_ = pc[k].y + t * (pc[j].y - pc[k].y);
                                grow(0.5f * W + px_ / zn * sx, 0.5f * H - py_ / zn * sy);
                            }
                    if (nfront > 0) {
                        rect = make_float4(fmaxf(x0 - 1.f, -1.f), fmaxf(y0 - 1.f, -1.f), fminf(x1 + 1.f, W + 1.f), fminf(y1 + 1.f, H + 1.f));
                        diag = make_float4(u0 - 2.f, v0 - 2.f, u1 + 2.f, v1 + 2.f);
                    }
                    zmin = fmaxf(zmin, 0.f);
                }
                float4 *rr = rrect + (((size_t)env * ncam + ci) * m.ngeom + g) * 3;
                rr[0] = rect; rr[1] = diag; rr[2] = make_float4(zmin, 0.f, 0.f, 0.f);
            }
        }
        __syncwarp();
    }
}

#define AV_KDOP 13
__constant__ float av_kdop_dir[AV_KDOP][3] = {
    {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
    {0.70710678f, 0.70710678f, 0}, {0.70710678f, -0.70710678f, 0}, {0.70710678f, 0, 0.70710678f}, {0.70710678f, 0, -0.70710678f},
    {0, 0.70710678f, 0.70710678f}, {0, 0.70710678f, -0.70710678f},
    {0.57735027f, 0.57735027f, 0.57735027f}, {0.57735027f, 0.57735027f, -0.57735027f}, {0.57735027f, -0.57735027f, 0.57735027f},
    {0.57735027f, -0.57735027f, -0.57735027f}};

struct RGeom {
    float pos[3], mat[9], size[3], rgb[3];
    float ol[3];            // camera centre in the geom frame (every primary ray of the tile starts there)
    float kd[AV_KDOP][2];   // slab intervals RELATIVE to ol: mesh = the hull's 26-DOP, box = its 3 axis slabs
    float rect[4];          // screen rectangle of the geom in this camera (pixel-level cull before the ray test)
    float diag[4];          // and its extents along x+y, x-y (lo, lo, hi, hi): a screen-space 8-DOP
    float zmin;             // nearest depth of its bounding box: no hit can be closer (t >= depth >= zmin)
    int type, gid, tex;
};

// one slab [a, b] (relative to the ray origin) along a direction whose ray component is dd: entry/exit update, branch-free.
// 1/dd is the bare MUFU reciprocal: a ray parallel to the slab gets +-inf, so both bounds land on the same side when the
// origin is outside (miss) and on opposite sides when it is inside (no constraint).  The entering slab and face are carried
// in the 5 low mantissa bits of the entry distance (slab << 1 | dd < 0), so the running maximum is one FMNMX instead of a
// compare and three selects; the price is 2^-19 relative on t (2 um at 1 m).
__device__ __forceinline__ float av_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define AV_SLAB(K, DD)                                                                                              \
    {                                                                                                               \
        float dd_ = (DD), inv_ = av_rcp(dd_);                                                                       \
        float t0_ = g.kd[K][0] * inv_, t1_ = g.kd[K][1] * inv_;                                                     \
        float lo_ = fminf(fminf(t0_, t1_), 3e38f), hi_ = fmaxf(t0_, t1_);   /* +inf would encode to a NaN */          \
        lo_ = __int_as_float((__float_as_int(lo_) & ~31) | (2 * (K)) | (int)((unsigned int)__float_as_int(dd_) >> 31)); \
        tn = fmaxf(tn, lo_);                                                                                        \
        tf = fminf(tf, hi_);                                                                                        \
    }

// nearest intersection of the ray ol + t dl (both in the geom frame, t > 1e-4) with one geom; returns t (or 1e30) and the
// normal IN THE GEOM FRAME (the caller rotates only the winner's)
__device__ __forceinline__ float ray_geom(const RGeom &g, V3 dl, V3 &nl) {
    const float INF = 1e30f;
    V3 ol = v3(g.ol[0], g.ol[1], g.ol[2]);
    if (g.type == AV_GEOM_MESH || g.type == AV_GEOM_BOX) {
        float tn = -INF, tf = INF;
        AV_SLAB(0, dl.x) AV_SLAB(1, dl.y) AV_SLAB(2, dl.z)
        if (tn > tf || tf <= 1e-4f) return INF;
        if (g.type == AV_GEOM_MESH) {
            const float h = 0.70710678f, q = 0.57735027f;
            AV_SLAB(3, h * (dl.x + dl.y)) AV_SLAB(4, h * (dl.x - dl.y)) AV_SLAB(5, h * (dl.x + dl.z)) AV_SLAB(6, h * (dl.x - dl.z))
            AV_SLAB(7, h * (dl.y + dl.z)) AV_SLAB(8, h * (dl.y - dl.z))
            AV_SLAB(9, q * (dl.x + dl.y + dl.z)) AV_SLAB(10, q * (dl.x + dl.y - dl.z)) AV_SLAB(11, q * (dl.x - dl.y + dl.z))
            AV_SLAB(12, q * (dl.x - dl.y - dl.z))
            if (tn > tf) return INF;
        }
        if (tn <= 1e-4f) return INF;
        const int code = __float_as_int(tn) & 31, ax = code >> 1;   // entered through the low face when dd > 0: normal = -dir
        const float sg = (code & 1) ? 1.f : -1.f;
        nl = v3(av_kdop_dir[ax][0], av_kdop_dir[ax][1], av_kdop_dir[ax][2]) * sg;
        return tn;
    }
    if (g.type == AV_GEOM_SPHERE) {
        float r = g.size[0], b = dot(ol, dl), cc = dot(ol, ol) - r * r, disc = b * b - cc;
        if (disc < 0.f) return INF;
        float t = -b - sqrtf(disc);
        if (t <= 1e-4f) return INF;
        nl = normalized(ol + dl * t);
        return t;
    }
    // cylinder along the local z axis
    float r = g.size[0], hh = g.size[1], best = INF;
    float a = dl.x * dl.x + dl.y * dl.y, b = ol.x * dl.x + ol.y * dl.y, cc = ol.x * ol.x + ol.y * ol.y - r * r;
    if (a > 1e-12f) {
        float disc = b * b - a * cc;
        if (disc >= 0.f) {
            float t = (-b - sqrtf(disc)) / a, z = ol.z + t * dl.z;
            if (t > 1e-4f && fabsf(z) <= hh) { best = t; nl = normalized(v3(ol.x + t * dl.x, ol.y + t * dl.y, 0.f)); }
        }
    }
    if (fabsf(dl.z) > 1e-12f) {
        float s = dl.z > 0.f ? -1.f : 1.f, t = (s * hh - ol.z) / dl.z;
        float x = ol.x + t * dl.x, y = ol.y + t * dl.y;
        if (t > 1e-4f && t < best && x * x + y * y <= r * r) { best = t; nl = v3(0.f, 0.f, s); }
    }
    return best;
}

__global__ void __launch_bounds__(AV_RT_THREADS) avsim_render_kernel(const __grid_constant__ DevModel m, const float *__restrict__ rpose,
                                                                     const float4 *__restrict__ rrect, const float *__restrict__ geom_rgb, const int *__restrict__ geom_visible,
                                                                     const float *__restrict__ cam_fovy, const int *__restrict__ cam_ids, int ncam,
                                                                     int ncam_all, int H, int W, unsigned char *__restrict__ dst, int id_mode) {
    __shared__ RGeom sg[AV_RT_MAXG];
    __shared__ int s_n;
    __shared__ int s_gid[AV_NG];
    __shared__ float s_z[AV_NG];
    __shared__ unsigned int s_tile[AV_RT_H][AV_RT_W * 3 / 4];
    const int tiles_x = (W + AV_RT_W - 1) / AV_RT_W;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x, ci = blockIdx.y, env = blockIdx.z;
    const int tid = threadIdx.x, wq = tid >> 5, ln = tid & 31;
    const int cam = cam_ids[ci];
    const float *base = rpose + (size_t)env * (m.ngeom + ncam_all) * 12;
    const float *cp = base + 12 * (m.ngeom + cam);
    V3 o = ld3(cp);
    M3 Rc = ldm3(cp + 3);
    const float th = tanf(0.5f * cam_fovy[cam] * 0.017453292519943295f), aspect = (float)W / (float)H;
    if (tid == 0) s_n = 0;
    __syncthreads();
    // region cull: the geom's screen-space 8-DOP (prep kernel) against this 64 x 32 region; the kept geoms are ordered front
    // to back by the nearest depth of their boxes so a pixel block can stop at the first candidate that starts behind its hits;
    // then all threads fill the records word by word.  This set-up is paid once per 2048 pixels.
    {
        const float x0 = (float)(tx * AV_RT_W), y0 = (float)(ty * AV_RT_H), x1 = x0 + AV_RT_W, y1 = y0 + AV_RT_H;
        const float4 *rr = rrect + ((size_t)env * ncam + ci) * m.ngeom * 3;
        for (int g = tid; g < m.ngeom; g += blockDim.x) {
            float4 q = rr[3 * g];
            if (q.x > q.z || q.x >= x1 || q.z < x0 || q.y >= y1 || q.w < y0) continue;
            float4 dg = rr[3 * g + 1];   // region extents along x+y: [x0+y0, x1+y1], along x-y: [x0-y1, x1-y0]
            if (dg.x > x1 + y1 || dg.z < x0 + y0 || dg.y > x1 - y0 || dg.w < x0 - y1) continue;
            int k = atomicAdd(&s_n, 1);
            if (k < AV_NG) { s_gid[k] = g; s_z[k] = rr[3 * g + 2].x; }
        }
        __syncthreads();
        const int nall = min(s_n, AV_NG);          // every geom that overlaps the region (a model has at most AV_NG)
        const int nk = min(nall, AV_RT_MAXG);      // ... of which the AV_RT_MAXG nearest are kept
        int my_g = 0, my_rank = AV_NG;
        if (tid < nall) {   // rank sort: ties broken by geom id, so the order, the kept set and the picture do not depend on the atomics
            my_g = s_gid[tid];
            my_rank = 0;
            float z = s_z[tid];
            for (int j = 0; j < nall; j++) my_rank += (s_z[j] < z || (s_z[j] == z && s_gid[j] < my_g)) ? 1 : 0;
        }
        __syncthreads();
        if (my_rank < nk) s_gid[my_rank] = my_g;
        __syncthreads();
        for (int idx = tid; idx < nk * 64; idx += blockDim.x) {
            const int k = idx >> 6, f = idx & 63, g = s_gid[k];
            RGeom &r = sg[k];
            const int ty_ = m.geom_type[g];
            if (f < 3) r.pos[f] = base[12 * g + f];
            else if (f < 12) r.mat[f - 3] = base[12 * g + f];
            else if (f < 15) r.size[f - 12] = ty_ == AV_GEOM_MESH ? m.geom_aabb[3 * g + f - 12] : m.geom_size[3 * g + f - 12];
            else if (f < 18) r.rgb[f - 15] = geom_rgb[4 * g + f - 15];
            else if (f < 18 + 2 * AV_KDOP) {
                if (ty_ == AV_GEOM_MESH) (&r.kd[0][0])[f - 18] = m.hull_kdop[(size_t)m.geom_hull[g] * AV_KDOP * 2 + f - 18];
                else if (f < 24) (&r.kd[0][0])[f - 18] = ((f & 1) ? 1.f : -1.f) * m.geom_size[3 * g + (f - 18) / 2];   // box: +-half size
            } else if (f < 48) r.rect[f - 44] = (&rr[3 * g].x)[f - 44];
            else if (f == 48) r.type = ty_;
            else if (f == 49) r.gid = g;
            else if (f == 50) r.tex = m.geom_tex[g];
            else if (f < 54) {   // camera centre in the geom frame: column f-51 of R dotted with (o - c)
                const float *gp = base + 12 * g;
                const int i = f - 51;
                r.ol[i] = gp[3 + i] * (o.x - gp[0]) + gp[6 + i] * (o.y - gp[1]) + gp[9 + i] * (o.z - gp[2]);
            } else if (f < 58) r.diag[f - 54] = (&rr[3 * g + 1].x)[f - 54];
            else if (f == 58) r.zmin = rr[3 * g + 2].x;
        }
        __syncthreads();
        for (int idx = tid; idx < nk * 32; idx += blockDim.x) {   // slabs relative to the ray origin
            const int k = idx >> 5, f = idx & 31;
            RGeom &r = sg[k];
            if (f < 2 * AV_KDOP) {
                const int sl = f >> 1;
                (&r.kd[0][0])[f] -= r.ol[0] * av_kdop_dir[sl][0] + r.ol[1] * av_kdop_dir[sl][1] + r.ol[2] * av_kdop_dir[sl][2];
            }
        }
    }
    __syncthreads();
    const int ng = min(s_n, AV_RT_MAXG);
    // each warp shades the 8 x 4 pixel blocks of one 8-pixel column strip of the region.  The lanes keep the screen extents of
    // "their" one or two candidates in registers (x test once per strip); per block a ballot of the y / diagonal tests gives the
    // candidates that overlap these 32 pixels, so the ray loop visits only those
    static_assert(AV_RT_W / 8 == AV_RT_THREADS / 32, "one column strip per warp");
    const int sbx = wq, lx = sbx * 8 + (ln & 7), px = tx * AV_RT_W + lx;
    const float bx0 = (float)(tx * AV_RT_W + sbx * 8), bx1 = bx0 + 8.f;
    float cr[AV_RT_MAXG / 32][8];
    bool xin[AV_RT_MAXG / 32];
#pragma unroll
    for (int h = 0; h < AV_RT_MAXG / 32; h++) {
        const int k = ln + 32 * h;
        xin[h] = false;
        if (k < ng) {
            const RGeom &G = sg[k];
#pragma unroll
            for (int i = 0; i < 4; i++) { cr[h][i] = G.rect[i]; cr[h][4 + i] = G.diag[i]; }
            xin[h] = !(cr[h][0] >= bx1 || cr[h][2] < bx0);
        }
    }
    // ray direction = normalize(Rc (x, y, -1)): the x part is fixed per lane, only y changes from block to block
    const float inv_w = 2.f / W, inv_h = 2.f / H;
    const float rx = ((px + 0.5f) * inv_w - 1.f) * th * aspect;
    const V3 dx = v3(Rc.m[0] * rx - Rc.m[2], Rc.m[3] * rx - Rc.m[5], Rc.m[6] * rx - Rc.m[8]);
    for (int sby = 0; sby < AV_RT_H / 4; sby++) {
        const int ly = sby * 4 + (ln >> 3), py = ty * AV_RT_H + ly;
        unsigned int mask[AV_RT_MAXG / 32];
        {
            const float by0 = (float)(ty * AV_RT_H + sby * 4), by1 = by0 + 4.f;
#pragma unroll
            for (int h = 0; h < AV_RT_MAXG / 32; h++) {
                if (32 * h >= ng) { mask[h] = 0u; continue; }   // uniform: most regions keep fewer than 32 candidates
                const bool in = xin[h] && !(cr[h][1] >= by1 || cr[h][3] < by0 || cr[h][4] > bx1 + by1 || cr[h][6] < bx0 + by0 ||
                                            cr[h][5] > bx1 - by0 || cr[h][7] < bx0 - by1);
                mask[h] = __ballot_sync(0xffffffffu, in);
            }
        }
        V3 d;
        {
            const float ry = (1.f - (py + 0.5f) * inv_h) * th;
            const V3 du = v3(dx.x + Rc.m[1] * ry, dx.y + Rc.m[4] * ry, dx.z + Rc.m[7] * ry);
            d = du * rsqrtf(dot(du, du));
        }
        float best = 1e30f;
        V3 bn = v3(0, 0, 1);
        int bi = -1;
        bool done = false;
#pragma unroll
        for (int h = 0; h < AV_RT_MAXG / 32; h++) {
            unsigned int mm = done ? 0u : mask[h];
            while (mm) {
                const int k = __ffs(mm) - 1 + 32 * h;
                mm &= mm - 1;
                const RGeom &G = sg[k];
                if (__all_sync(0xffffffffu, best <= G.zmin)) { done = true; break; }   // front-to-back: nothing further can be nearer
                V3 dl = v3(G.mat[0] * d.x + G.mat[3] * d.y + G.mat[6] * d.z, G.mat[1] * d.x + G.mat[4] * d.y + G.mat[7] * d.z,
                           G.mat[2] * d.x + G.mat[5] * d.y + G.mat[8] * d.z);
                V3 n;
                float t = ray_geom(G, dl, n);
                if (t < best) { best = t; bn = n; bi = k; }
            }
        }
        if (bi >= 0) {   // the winner's normal, geom frame -> world
            const float *Rm = sg[bi].mat;
            bn = v3(Rm[0] * bn.x + Rm[1] * bn.y + Rm[2] * bn.z, Rm[3] * bn.x + Rm[4] * bn.y + Rm[5] * bn.z, Rm[6] * bn.x + Rm[7] * bn.y + Rm[8] * bn.z);
        }
        float r, g, b;
        if (id_mode) {   // test hook: the index of the geom the primary ray hits (255: none) instead of a colour
            r = g = b = (bi >= 0 ? (float)sg[bi].gid : 255.f) / 255.f;
        } else if (bi >= 0) {
            const float *L = m.light;   // dir 3 | diffuse | headlight ambient | headlight diffuse
            float head = fmaxf(0.f, -dot(bn, d));
            float sun = fmaxf(0.f, -(bn.x * L[0] + bn.y * L[1] + bn.z * L[2]));
            float lum = L[4] + L[5] * head + L[3] * sun;
            float cr = sg[bi].rgb[0], cg = sg[bi].rgb[1], cb = sg[bi].rgb[2];
            if (sg[bi].tex && bn.z > 0.9f) {   // table top: planar map of the diffuse texture over the top face
                V3 hit = o + d * best;
                float u = (hit.x - sg[bi].pos[0]) / (2.f * sg[bi].size[0]) + 0.5f, v = (hit.y - sg[bi].pos[1]) / (2.f * sg[bi].size[1]) + 0.5f;
                int iu = min(127, max(0, (int)(u * 128.f))), iv = min(127, max(0, (int)((1.f - v) * 128.f)));
                const float *t = m.table_tex + 3 * (iv * 128 + iu);
                cr = t[0]; cg = t[1]; cb = t[2];
            }
            r = cr * lum; g = cg * lum; b = cb * lum;
        } else if (d.z < 0.f) {  // floor below the horizon
            r = 0.2f; g = 0.3f; b = 0.4f;
        } else {                 // gradient sky (scene.xml:34)
            float k = fminf(1.f, d.z * 1.5f);
            r = 0.3f * (1 - k); g = 0.5f * (1 - k); b = 0.7f * (1 - k);
        }
        unsigned char *bytes = reinterpret_cast<unsigned char *>(&s_tile[ly][0]);
        bytes[3 * lx + 0] = (unsigned char)(fminf(1.f, r) * 255.f + 0.5f);
        bytes[3 * lx + 1] = (unsigned char)(fminf(1.f, g) * 255.f + 0.5f);
        bytes[3 * lx + 2] = (unsigned char)(fminf(1.f, b) * 255.f + 0.5f);
    }
    __syncthreads();
    // write the region as 32-bit words: 48 words (192 bytes) per region row, contiguous in the image row
    const int words_row = AV_RT_W * 3 / 4;
    for (int idx = tid; idx < AV_RT_H * words_row; idx += AV_RT_THREADS) {
        int row = idx / words_row, w = idx % words_row, y = ty * AV_RT_H + row, x0 = tx * AV_RT_W;
        if (y < H && x0 * 3 + 4 * w + 3 < W * 3) {
            size_t off = ((((size_t)env * ncam + ci) * H + y) * W + x0) * 3 + 4 * (size_t)w;
            *reinterpret_cast<unsigned int *>(dst + off) = s_tile[row][w];
        }
    }
}
