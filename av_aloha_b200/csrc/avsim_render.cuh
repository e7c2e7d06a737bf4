// avsim_render.cuh -- K9: multi-camera renderer, writes uint8 [B][ncam][H][W][3] (what physics.render(h, w, camera_id)
// returns per camera in get_obs, reference gym_guided_vision/env.py:180-188,195-200).
//
// Fidelity (stated, not hidden): this is a ray caster over the PHYSICS geoms, not a rasteriser of the 588 k visual
// triangles.  Task objects, table and finger geometry are drawn exactly as their primitives; every mesh geom (robot
// links, frame extrusions) is drawn as the oriented bounding box of its convex hull, in the colour of the visual mesh it
// stands for.  Flat colours (no table texture), MuJoCo-style headlight (ambient 0.3 + diffuse 0.6, scene.xml:9) plus one
// fixed directional light, no shadows, gradient background.  The reference's own renders are declared non-deterministic
// (gym_guided_vision/__init__.py:92-94) and no GL context exists here, so there is no pixel oracle: tests check camera
// geometry (a known point projects to the expected pixel), determinism and coverage, not image equality.
//
// Two kernels:
//   avsim_render_prep_kernel  warp per environment: forward kinematics from qpos, world pose of every geom and camera
//                             -> rpose[B][ngeom + ncam_all][12] (pos 3 | rotation 9), and the screen rectangle of every
//                             geom's oriented box in every requested camera -> rrect[B][ncam][ngeom]; L2 resident
//   avsim_render_kernel       block per (32 x 8 pixel tile, camera, environment): culls geoms by screen rectangle
//                             into shared memory, one primary ray per thread, nearest hit, shade, stage the tile in shared
//                             memory and write it as 32-bit words (96 contiguous bytes per tile row).  HBM-write bound:
//                             H*W*3 bytes per image.
#pragma once
#include "avsim_step.cuh"

#define AV_RT_W 32
#define AV_RT_H 8
#define AV_RT_MAXG 48   // candidate geoms per tile

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
                float4 rect = make_float4(1.f, 1.f, 0.f, 0.f);
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
                    float x0 = 1e30f, y0 = 1e30f, x1 = -1e30f, y1 = -1e30f;
                    V3 pc[8];
                    int nfront = 0;
                    for (int k = 0; k < 8; k++) {
                        V3 pw = c + mul(R, v3((k & 1) ? h.x : -h.x, (k & 2) ? h.y : -h.y, (k & 4) ? h.z : -h.z));
                        pc[k] = mulT(Rc, pw - o);
                        if (pc[k].z <= -zn) {
                            nfront++;
                            float iz = -1.0f / pc[k].z, x = 0.5f * W + pc[k].x * iz * sx, y = 0.5f * H - pc[k].y * iz * sy;
                            x0 = fminf(x0, x); x1 = fmaxf(x1, x); y0 = fminf(y0, y); y1 = fmaxf(y1, y);
                        }
                    }
                    if (nfront > 0 && nfront < 8)
                        for (int k = 0; k < 8; k++)
                            for (int bit = 1; bit < 8; bit <<= 1) {
                                int j = k ^ bit;
                                if (j < k || (pc[k].z <= -zn) == (pc[j].z <= -zn)) continue;
                                float t = (-zn - pc[k].z) / (pc[j].z - pc[k].z);
                                float px_ = pc[k].x + t * (pc[j].x - pc[k].x), py_ = pc[k].y + t * (pc[j].y - pc[k].y);
                                float x = 0.5f * W + px_ / zn * sx, y = 0.5f * H - py_ / zn * sy;
                                x0 = fminf(x0, x); x1 = fmaxf(x1, x); y0 = fminf(y0, y); y1 = fmaxf(y1, y);
                            }
                    if (nfront > 0) rect = make_float4(fmaxf(x0 - 1.f, -1.f), fmaxf(y0 - 1.f, -1.f), fminf(x1 + 1.f, W + 1.f), fminf(y1 + 1.f, H + 1.f));
                }
                rrect[((size_t)env * ncam + ci) * m.ngeom + g] = rect;
            }
        }
        __syncwarp();
    }
}

struct RGeom {
    float pos[3], mat[9], size[3], rgb[3];
    int type;
};

// nearest intersection of the ray o + t d (t > tmin) with one geom; returns t (or 1e30) and the world normal
__device__ inline float ray_geom(const RGeom &g, V3 o, V3 d, V3 &nrm) {
    M3 R;
#pragma unroll
    for (int i = 0; i < 9; i++) R.m[i] = g.mat[i];
    V3 c = v3(g.pos[0], g.pos[1], g.pos[2]);
    V3 ol = mulT(R, o - c), dl = mulT(R, d);
    const float INF = 1e30f;
    if (g.type == AV_GEOM_SPHERE) {
        float r = g.size[0], b = dot(ol, dl), cc = dot(ol, ol) - r * r, disc = b * b - cc;
        if (disc < 0.f) return INF;
        float t = -b - sqrtf(disc);
        if (t <= 1e-4f) return INF;
        nrm = mul(R, normalized(ol + dl * t));
        return t;
    }
    if (g.type == AV_GEOM_CYLINDER) {
        float r = g.size[0], hh = g.size[1], best = INF;
        float a = dl.x * dl.x + dl.y * dl.y, b = ol.x * dl.x + ol.y * dl.y, cc = ol.x * ol.x + ol.y * ol.y - r * r;
        if (a > 1e-12f) {
            float disc = b * b - a * cc;
            if (disc >= 0.f) {
                float t = (-b - sqrtf(disc)) / a, z = ol.z + t * dl.z;
                if (t > 1e-4f && fabsf(z) <= hh) { best = t; nrm = mul(R, normalized(v3(ol.x + t * dl.x, ol.y + t * dl.y, 0.f))); }
            }
        }
        if (fabsf(dl.z) > 1e-12f) {
            float s = dl.z > 0.f ? -1.f : 1.f, t = (s * hh - ol.z) / dl.z;
            float x = ol.x + t * dl.x, y = ol.y + t * dl.y;
            if (t > 1e-4f && t < best && x * x + y * y <= r * r) { best = t; nrm = mul(R, v3(0.f, 0.f, s)); }
        }
        return best;
    }
    // box, and the hull's oriented bounding box for meshes: slab test in the local frame
    float tn = -INF, tf = INF;
    int ax = 0;
    float sg = 1.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        float ok = comp(ol, k), dk = comp(dl, k), h = g.size[k];
        if (fabsf(dk) < 1e-12f) {
            if (fabsf(ok) > h) return INF;
            continue;
        }
        float inv = 1.0f / dk, t0 = (-h - ok) * inv, t1 = (h - ok) * inv;
        float s0 = -1.f;
        if (t0 > t1) { float tmp = t0; t0 = t1; t1 = tmp; s0 = 1.f; }
        if (t0 > tn) { tn = t0; ax = k; sg = s0; }
        tf = fminf(tf, t1);
    }
    if (tn > tf || tn <= 1e-4f) return INF;
    nrm = colm(R, ax) * sg;
    return tn;
}

__global__ void __launch_bounds__(AV_RT_W *AV_RT_H) avsim_render_kernel(const __grid_constant__ DevModel m, const float *__restrict__ rpose,
                                                                        const float4 *__restrict__ rrect, const float *__restrict__ geom_rgb, const int *__restrict__ geom_visible,
                                                                        const float *__restrict__ cam_fovy, const int *__restrict__ cam_ids, int ncam,
                                                                        int ncam_all, int H, int W, unsigned char *__restrict__ dst) {
    __shared__ RGeom sg[AV_RT_MAXG];
    __shared__ int s_n;
    __shared__ unsigned int s_tile[AV_RT_H][AV_RT_W * 3 / 4];
    const int tiles_x = (W + AV_RT_W - 1) / AV_RT_W;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x, ci = blockIdx.y, env = blockIdx.z;
    const int tid = threadIdx.x, lx = tid % AV_RT_W, ly = tid / AV_RT_W;
    const int cam = cam_ids[ci];
    const float *base = rpose + (size_t)env * (m.ngeom + ncam_all) * 12;
    const float *cp = base + 12 * (m.ngeom + cam);
    V3 o = ld3(cp);
    M3 Rc = ldm3(cp + 3);
    const float th = tanf(0.5f * cam_fovy[cam] * 0.017453292519943295f), aspect = (float)W / (float)H;
    auto ray = [&](float px, float py) {
        float x = (2.f * px / W - 1.f) * th * aspect, y = (1.f - 2.f * py / H) * th;
        return normalized(mul(Rc, v3(x, y, -1.f)));
    };
    if (tid == 0) s_n = 0;
    __syncthreads();
    // cull: the geom's screen rectangle (prep kernel) against this tile
    {
        const float x0 = (float)(tx * AV_RT_W), y0 = (float)(ty * AV_RT_H);
        const float4 *rr = rrect + ((size_t)env * ncam + ci) * m.ngeom;
        for (int g = tid; g < m.ngeom; g += blockDim.x) {
            float4 q = rr[g];
            if (q.x > q.z || q.x >= x0 + AV_RT_W || q.z < x0 || q.y >= y0 + AV_RT_H || q.w < y0) continue;
            int k = atomicAdd(&s_n, 1);
            if (k < AV_RT_MAXG) {
                RGeom &r = sg[k];
                for (int i = 0; i < 3; i++) r.pos[i] = base[12 * g + i];
                for (int i = 0; i < 9; i++) r.mat[i] = base[12 * g + 3 + i];
                int ty_ = m.geom_type[g];
                r.type = ty_;
                for (int i = 0; i < 3; i++) {
                    r.size[i] = ty_ == AV_GEOM_MESH ? m.geom_aabb[3 * g + i] : m.geom_size[3 * g + i];
                    r.rgb[i] = geom_rgb[4 * g + i];
                }
            }
        }
    }
    __syncthreads();
    const int ng = min(s_n, AV_RT_MAXG);
    const int px = tx * AV_RT_W + lx, py = ty * AV_RT_H + ly;
    V3 d = ray(px + 0.5f, py + 0.5f);
    float best = 1e30f;
    V3 bn = v3(0, 0, 1);
    int bi = -1;
    for (int k = 0; k < ng; k++) {
        V3 n;
        float t = ray_geom(sg[k], o, d, n);
        if (t < best) { best = t; bn = n; bi = k; }
    }
    float r, g, b;
    if (bi >= 0) {
        float head = fmaxf(0.f, -dot(bn, d));
        float sun = fmaxf(0.f, dot(bn, normalized(v3(0.3f, -0.2f, 1.f))));
        float lum = 0.3f + 0.6f * head + 0.15f * sun;
        r = sg[bi].rgb[0] * lum; g = sg[bi].rgb[1] * lum; b = sg[bi].rgb[2] * lum;
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
    __syncthreads();
    // write the tile as 32-bit words: 24 words (96 bytes) per tile row, contiguous in the image row
    const int words_row = AV_RT_W * 3 / 4;
    if (tid < AV_RT_H * words_row) {
        int row = tid / words_row, w = tid % words_row, y = ty * AV_RT_H + row, x0 = tx * AV_RT_W;
        if (y < H && x0 * 3 + 4 * w + 3 < W * 3) {
            size_t off = ((((size_t)env * ncam + ci) * H + y) * W + x0) * 3 + 4 * (size_t)w;
            *reinterpret_cast<unsigned int *>(dst + off) = s_tile[row][w];
        }
    }
}
