// avsim_step.cuh -- the per-substep pipeline, one warp per environment (sm_100a, fp32).
//
// Replaces, stage by stage, what `physics.step(nstep=20)` (reference gym_guided_vision/gym_guided_vision/env.py:218)
// runs inside MuJoCo's mj_step [third-party, not in the reference tree; semantics per SURVEY.md Appendix A]:
//   K2 kinematics/comPos  -> stage_kinematics      K5 makeConstraint -> stage_rows
//   K3 crb/factorM/rne    -> stage_inertia/smooth  K6 solver+noslip  -> stage_solve
//   K4 collision          -> stage_collision       K7 Euler          -> stage_integrate
//   K8 reward/agent_pos   -> stage_reward (env.py:425-472,546-589,640-690,738-779,820-863; 169-178)
// All per-environment working state lives in the warp's shared-memory slice (EnvS); only the contact Jacobian
// blocks go to an L2-resident global scratch.  Lane roles change per stage: lane = kinematic tree for the
// recursive passes, lane = dof for joint-space vectors, lane = candidate pair for primitive narrowphase,
// lane = (row-half, column) of a 6x16 Jacobian block in the solver.  No __syncthreads: a block is one warp.
#pragma once
#include "avsim_collide.cuh"

#define AV_CB_GEO (52 + 6 * AV_JW)   // offset of the contact geometry inside a scratch block (layout: avsim_solve.cuh)
static_assert((AV_CB_GEO * 4) % 16 == 0 && (AV_CBLK * 4) % 16 == 0, "bulk copies move 16-byte multiples");

// ---- optional per-stage cycle counters (-DAVSIM_PROFILE; read back with avsim_stage_cycles)
enum { PF_LOAD = 0, PF_KIN, PF_INERTIA, PF_BROAD, PF_PRIM, PF_CONVEX, PF_SMOOTH, PF_ROWS_S, PF_ROWS_C, PF_SOLVE, PF_INTEGRATE,
       PF_OUT, PF_NW_INIT, PF_NW_GRAD, PF_NW_HESS, PF_NW_CHOL, PF_NW_LS, PF_NW_WAIT, PF_N };
#ifdef AVSIM_PROFILE
__device__ unsigned long long g_prof[PF_N];
struct Prof {
    long long t;
    __device__ __forceinline__ void start() { t = clock64(); }
    __device__ __forceinline__ void mark(int k, int lane) {
        long long n = clock64();
        if (lane == 0) atomicAdd(&g_prof[k], (unsigned long long)(n - t));
        t = n;
    }
};
#else
struct Prof {
    __device__ __forceinline__ void start() {}
    __device__ __forceinline__ void mark(int, int) {}
};
#endif

// Shared-memory slice of one environment.  Arrays whose lifetimes do not overlap inside a substep share storage (the
// slice size sets how many environments fit on an SM, and the kernel is latency-bound: more resident warps = more
// throughput).  Stage order: kinematics -> inertia -> collision -> smooth -> rows_scalar -> rows_contact -> solve ->
// integrate.  Lifetimes:
//   crb   : composite inertias (inertia) | cacc,cfrc (smooth) | Cholesky factor L (end of inertia; integrate) |
//           J staging of the contact being assembled + M^-1 J^T of the scalar rows (rows_scalar..solve)
//   u1    : broadphase candidates + world AABBs (kinematics..collision) | cvel, cdofdot (smooth) | contact forces, multipliers
//           (rows_contact..solve)
//   u2    : cinert, xipos (kinematics..smooth) | solver schedule (solve_begin..solve)
//   u1+u2 : the Newton solver's Hessian (stage_newton only)
//   u3    : geom centres (kinematics..collision) | scalar constraint rows (rows_scalar..solve)
struct EnvS {
    // ---- head: everything that crosses a kernel boundary of the split pipeline (substep kernel -> solver kernel -> substep
    // kernel; avsim_kernels.cuh).  One contiguous, 16-byte aligned record of AV_HEAD_FLOATS floats that is moved between the
    // per-environment image in global memory (BatchState::heads, L2 resident) and shared memory by one bulk copy each way.
    float qpos[AV_NQ], qvel[AV_NVP], ctrl[24], warm[AV_NVP];
    float M[AV_MPK], Minv[AV_MBLK];   // M: packed lower triangles (av_mtri), read by the two factorisations and one mat-vec per substep
    float qfrc_smooth[AV_NVP], qacc_smooth[AV_NVP], acc[AV_NVP], qfrc_bias[AV_NVP];
    union {
        float crb[AV_NB * 12];
        float L[AV_MBLK];
        struct {                                      // rows_scalar .. solve: the region is free between smooth and integrate
            __align__(16) float stage[2 * 6 * AV_JW];  // J | MinvJT of the contact being assembled
            float sc_MJ[AV_NSC * AV_TD];
        };
    };
    union {
        float gpos[AV_NG * 3];  // world centre of every geom
        struct {
            int sc_dof1[AV_NSC], sc_dof2[AV_NSC], sc_tree[AV_NSC], sc_key[AV_NSC];
            float sc_c1[AV_NSC], sc_c2[AV_NSC], sc_b[AV_NSC], sc_R[AV_NSC], sc_f[AV_NSC], sc_lo[AV_NSC], sc_hi[AV_NSC],
                sc_A[AV_NSC], sc_aref[AV_NSC];
        };
    };
    // contacts (collision .. outputs); position / frame / distance / friction live in the contact's global scratch block
    int c_info[AV_NCON];  // geom1 | geom2 << 8 | dim << 16 | excluded << 20 | tree1 << 21 | tree2 << 24 (7: none; set by the row assembly)
    int tree_pk[AV_NTREE + 2];   // per kinematic tree: dofadr | dofnum << 6 (copy of the model table; the contact's trees sit in c_info)
    int ncon, nsc, ncand_p, ncand_c, nkeep, nslot, status, pad_;
    // ---- solver scratch: also resident in the solver kernel's (smaller) slices
    union {
        struct {
            union {
                struct { float gaabb[AV_NG * 3]; int cand_p[AV_NCAND], cand_c[AV_NCAND], keep_c[AV_NKEEP]; };
                struct { float cvel[AV_NB * 6], cdofdot[AV_NV * 6]; };
                struct { float c_f[AV_NCON * 6], c_lam[AV_NCON]; };   // contact forces and cone multipliers (rows_contact..solve, cache store)
            };
            union {
                struct { float cinert[AV_NB * 10], xipos[AV_NB * 3]; };
                unsigned short c_slot[AV_NCON];  // solver schedule: contact a | contact b << 8 (0xff: none) per half-warp slot
            };
        };
        // Newton solve (avsim_newton.cuh): Hessian / its Cholesky factor as a packed lower triangle.  Alive from the end of the row
        // assembly until the solver publishes its forces into c_f; the sweep schedule (c_slot) is built after that.
        float H[AV_NHP];
    };
    // ---- substep kernel only (recomputed every substep)
    float xpos[AV_NB * 3], xquat[AV_NB * 4];
    float torig[AV_NTREE * 3];
    float cdof[AV_NV * 6];
    // optional solver pipeline (-DAV_BULK_PREFETCH=1): double buffer for the contact block being updated / prefetched by
    // bulk async copies, its two mbarriers and how often each buffer has been filled (phase parity)
#if AV_BULK_PREFETCH
    __align__(16) float cbuf[2][AV_CB_GEO];
    unsigned long long mbar[2];
    unsigned cuse[2];
#endif
};
#define AV_HEAD_FLOATS ((int)(offsetof(EnvS, H) / 4))
// what the solver kernel stores back: the head + the contact forces / cone multipliers (c_f, c_lam: the first part of the solver scratch)
#define AV_HEADX_FLOATS (AV_HEAD_FLOATS + AV_NCON * 7)
#define AV_SOLVER_SLICE_BYTES (offsetof(EnvS, xpos))
static_assert((AV_HEAD_FLOATS * 4) % 16 == 0 && (AV_HEADX_FLOATS * 4) % 16 == 0 && AV_SOLVER_SLICE_BYTES % 16 == 0, "bulk copies move 16-byte multiples");
static_assert(sizeof(float[AV_NB * 12]) >= sizeof(float[AV_MBLK]), "L must fit inside crb");
static_assert(sizeof(float[AV_NB * 12]) >= sizeof(float[2 * 6 * AV_JW + AV_NSC * AV_TD]), "row staging must fit inside crb");
static_assert(AV_NHP <= AV_NG * 3 + 2 * AV_NCAND + AV_NKEEP + AV_NB * 13, "the Hessian must fit inside the two unions it overlays");
static_assert(sizeof(EnvS) % 16 == 0, "slices must keep 16-byte alignment");

// per-environment view of the constraint-force cache in global memory (BatchState.fc_*)
struct FCache {
    int *key, *n;
    float *val;
    int mode;
};
// identity of a contact: geom pair + ordinal among the (consecutive) contacts of that pair; same as the oracle's key
__device__ __forceinline__ int contact_key(const EnvS &S, int c);

__device__ __forceinline__ M3 body_mat(const EnvS &S, int b) { return q2m(ldq(S.xquat + 4 * b)); }

__device__ __forceinline__ int body_mask_has(const DevModel &m, int body, int dof) {
    // is `dof` on the path from the tree root to `body`?
    for (int i = m.body_lastdof[body]; i >= 0; i = m.dof_parent[i])
        if (i == dof) return 1;
    return 0;
}

// ------------------------------------------------------------------ K2: kinematics, lane = kinematic tree
__device__ AV_STAGE void stage_kinematics(const DevModel &m, EnvS &S, int lane) {
    if (lane < m.ntree) {
        int b0 = m.tree_bodyadr[lane], nb = m.tree_bodynum[lane];
        V3 org = v3(0, 0, 0);
        for (int b = b0; b < b0 + nb; b++) {
            int p = m.body_parent[b];
            M3 Rp = body_mat(S, p);
            V3 pos = ld3(S.xpos + 3 * p) + mul(Rp, ld3(m.body_pos + 3 * b));
            Q4 quat = qmul(ldq(S.xquat + 4 * p), ldq(m.body_quat + 4 * b));
            int j0 = m.body_jntadr[b], nj = m.body_jntnum[b];
            V3 anchor[2], axis[2];
            for (int k = 0; k < nj; k++) {
                int j = j0 + k, qa = m.jnt_qposadr[j], jt = m.jnt_type[j];
                if (jt == AV_JNT_FREE) {
                    pos = ld3(S.qpos + qa);
                    quat = qnormalize(ldq(S.qpos + qa + 3));
                    continue;
                }
                M3 R = q2m(quat);
                V3 ax_l = ld3(m.jnt_axis + 3 * j), jp = ld3(m.jnt_pos + 3 * j);
                V3 ax = mul(R, ax_l), anc = pos + mul(R, jp);
                float dq = S.qpos[qa] - m.qpos0[qa];
                if (jt == AV_JNT_SLIDE) {
                    pos = pos + ax * dq;
                } else {
                    float s = sinf(0.5f * dq);
                    Q4 rq = {cosf(0.5f * dq), s * ax_l.x, s * ax_l.y, s * ax_l.z};
                    quat = qmul(quat, rq);
                    pos = anc - mul(q2m(quat), jp);
                }
                if (k < 2) { anchor[k] = anc; axis[k] = ax; }
            }
            quat = qnormalize(quat);
            M3 R = q2m(quat);
            st3(S.xpos + 3 * b, pos); stq(S.xquat + 4 * b, quat);
            if (b == b0) { org = pos; st3(S.torig + 3 * lane, org); }
            // dof axes about the tree origin
            int d0 = m.body_dofadr[b];
            if (nj == 1 && m.jnt_type[j0] == AV_JNT_FREE) {
                for (int k = 0; k < 3; k++) {
                    S6 c = {v3(0, 0, 0), v3(k == 0, k == 1, k == 2)};
                    st6(S.cdof + 6 * (d0 + k), c);
                    V3 ax = colm(R, k);
                    S6 r = {ax, cross(pos - org, ax)};
                    st6(S.cdof + 6 * (d0 + 3 + k), r);
                }
            } else {
                for (int k = 0; k < nj && k < 2; k++) {
                    S6 c;
                    if (m.jnt_type[j0 + k] == AV_JNT_SLIDE) { c.a = v3(0, 0, 0); c.l = axis[k]; }
                    else { c.a = axis[k]; c.l = cross(anchor[k] - org, axis[k]); }
                    st6(S.cdof + 6 * (d0 + k), c);
                }
            }
            // spatial inertia about the tree origin
            V3 xi = pos + mul(R, ld3(m.body_ipos + 3 * b));
            st3(S.xipos + 3 * b, xi);
            float mass = m.body_mass[b];
            const float *i6 = m.body_inertia + 6 * b;
            M3 Ib;
            Ib.m[0] = i6[0]; Ib.m[1] = i6[3]; Ib.m[2] = i6[4]; Ib.m[3] = i6[3]; Ib.m[4] = i6[1]; Ib.m[5] = i6[5];
            Ib.m[6] = i6[4]; Ib.m[7] = i6[5]; Ib.m[8] = i6[2];
            M3 Rt;
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rt.m[3 * r + c] = R.m[3 * c + r];
            M3 Iw = mul(mul(R, Ib), Rt);
            V3 c = xi - org;
            float cc = dot(c, c);
            float *I = S.cinert + 10 * b;
            I[0] = Iw.m[0] + mass * (cc - c.x * c.x); I[1] = Iw.m[4] + mass * (cc - c.y * c.y);
            I[2] = Iw.m[8] + mass * (cc - c.z * c.z);
            I[3] = Iw.m[1] - mass * c.x * c.y; I[4] = Iw.m[2] - mass * c.x * c.z; I[5] = Iw.m[5] - mass * c.y * c.z;
            I[6] = mass * c.x; I[7] = mass * c.y; I[8] = mass * c.z; I[9] = mass;
        }
    }
    __syncwarp();
    for (int g = lane; g < m.ngeom; g += 32) {
        if (m.geom_static[g]) {   // world-welded geoms: compile-time pose (re-written every substep: the slots are shared)
            st3(S.gpos + 3 * g, ld3(m.geom_xpos0 + 3 * g));
            st3(S.gaabb + 3 * g, ld3(m.geom_xaabb0 + 3 * g));
            continue;
        }
        int b = m.geom_body[g];
        M3 Rb = body_mat(S, b);
        st3(S.gpos + 3 * g, ld3(S.xpos + 3 * b) + mul(Rb, ld3(m.geom_pos + 3 * g)));
        M3 Rg = mul(Rb, ldm3(m.geom_mat + 9 * g));
        V3 h = ld3(m.geom_aabb + 3 * g);
        st3(S.gaabb + 3 * g, v3(fabsf(Rg.m[0]) * h.x + fabsf(Rg.m[1]) * h.y + fabsf(Rg.m[2]) * h.z,
                                fabsf(Rg.m[3]) * h.x + fabsf(Rg.m[4]) * h.y + fabsf(Rg.m[5]) * h.z,
                                fabsf(Rg.m[6]) * h.x + fabsf(Rg.m[7]) * h.y + fabsf(Rg.m[8]) * h.z));
    }
    __syncwarp();
}

// ------------------------------------------------------------------ K3a: CRB, factorisation, block inverse
__device__ __forceinline__ int av_mtri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }   // packed symmetric index
// Cholesky of the nt x nt block at A (packed lower triangle) into Lb (lower, stride AV_TD); returns false on a non-positive pivot
__device__ inline bool chol_block(const float *A, float *Lb, int nt, const float *diag_add, float h) {
    bool ok = true;
    for (int i = 0; i < nt; i++)
        for (int j = 0; j <= i; j++) {
            float s = A[i * (i + 1) / 2 + j];
            if (i == j && diag_add) s += h * diag_add[i];
            for (int k = 0; k < j; k++) s -= Lb[i * AV_TD + k] * Lb[j * AV_TD + k];
            if (i == j) {
                if (!(s > 0)) { ok = false; s = 1e-10f; }
                Lb[i * AV_TD + i] = sqrtf(s);
            } else
                Lb[i * AV_TD + j] = s / Lb[j * AV_TD + j];
        }
    return ok;
}
__device__ inline void chol_block_solve(const float *Lb, int nt, float *x) {
    for (int i = 0; i < nt; i++) {
        float s = x[i];
        for (int k = 0; k < i; k++) s -= Lb[i * AV_TD + k] * x[k];
        x[i] = s / Lb[i * AV_TD + i];
    }
    for (int i = nt - 1; i >= 0; i--) {
        float s = x[i];
        for (int k = i + 1; k < nt; k++) s -= Lb[k * AV_TD + i] * x[k];
        x[i] = s / Lb[i * AV_TD + i];
    }
}

__device__ AV_STAGE void stage_inertia(const DevModel &m, EnvS &S, int lane) {
    for (int i = lane; i < AV_MPK; i += 32) S.M[i] = 0.f;
    if (lane < m.ntree) {  // composite inertias, leaves to root
        int b0 = m.tree_bodyadr[lane], nb = m.tree_bodynum[lane];
        for (int b = b0; b < b0 + nb; b++)
            for (int k = 0; k < 10; k++) S.crb[12 * b + k] = S.cinert[10 * b + k];
        for (int b = b0 + nb - 1; b > b0; b--) {
            int p = m.body_parent[b];
            for (int k = 0; k < 10; k++) S.crb[12 * p + k] += S.crb[12 * b + k];
        }
    }
    __syncwarp();
    for (int i = lane; i < m.nv; i += 32) {  // lane = dof: one row of M up the ancestor chain
        int t = m.dof_tree[i], d0 = m.tree_dofadr[t];
        S6 f = inert_mul(S.crb + 12 * m.dof_body[i], ld6(S.cdof + 6 * i));
        float *Mb = S.M + t * AV_MTRI;
        for (int j = i; j >= 0; j = m.dof_parent[j]) {           // ancestors have smaller indices: (i, j) is in the lower triangle
            float v = dot6(ld6(S.cdof + 6 * j), f);
            if (j == i) v += m.dof_armature[i];
            Mb[av_mtri(i - d0, j - d0)] = v;
        }
    }
    __syncwarp();
    if (lane < m.ntree) {
        int nt = m.tree_dofnum[lane];
        if (!chol_block(S.M + lane * AV_MTRI, S.L + lane * AV_TD * AV_TD, nt, nullptr, 0.f)) S.status |= 1;
    }
    __syncwarp();
    for (int w = lane; w < m.nv; w += 32) {  // lane = (tree, column) of the block inverse
        int t = m.dof_tree[w], c = w - m.tree_dofadr[t], nt = m.tree_dofnum[t];
        float x[AV_TD];
        for (int k = 0; k < AV_TD; k++) x[k] = (k == c) ? 1.f : 0.f;
        chol_block_solve(S.L + t * AV_TD * AV_TD, nt, x);
        for (int k = 0; k < nt; k++) S.Minv[t * AV_TD * AV_TD + k * AV_TD + c] = x[k];
    }
    __syncwarp();
}

// ------------------------------------------------------------------ K4: collision
__device__ inline Shape load_shape(const DevModel &m, const EnvS &S, int g) {
    Shape s;
    s.type = m.geom_type[g];
    s.size = ld3(m.geom_size + 3 * g);
    s.vert = nullptr; s.nvert = 0;
    if (m.geom_static[g]) {
        s.pos = ld3(m.geom_xpos0 + 3 * g);
        s.mat = ldm3(m.geom_xmat0 + 9 * g);
    } else {
        int b = m.geom_body[g];
        s.pos = ld3(S.gpos + 3 * g);
        s.mat = mul(body_mat(S, b), ldm3(m.geom_mat + 9 * g));
    }
    if (s.type == AV_GEOM_MESH) {
        int h = m.geom_hull[g];
        s.vert = m.hull_vert + m.hull_adr[h];
        s.nvert = m.hull_num[h];
    }
    return s;
}

__device__ inline void add_contact(const DevModel &m, EnvS &S, float *scratch, int slot, int g1, int g2, float dist, V3 pos, V3 nrm) {
    V3 t1, t2;
    make_frame(nrm, t1, t2);
    float *geo = scratch + slot * AV_CBLK + AV_CB_GEO;
    st3(geo, pos);
    st3(geo + 3, nrm); st3(geo + 6, t1); st3(geo + 9, t2);
    geo[12] = dist;
    int dim = max(m.geom_condim[g1], m.geom_condim[g2]);
    for (int k = 0; k < 3; k++) geo[13 + k] = fmaxf(m.geom_friction[3 * g1 + k], m.geom_friction[3 * g2 + k]);
    float margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]), gap = fmaxf(m.geom_gap[g1], m.geom_gap[g2]);
    int excluded = !(dist < margin - gap);
    S.c_info[slot] = g1 | (g2 << 8) | (dim << 16) | (excluded << 20);
}

// primitive candidates (sphere / box pairs): one candidate per lane; drains S.cand_p
__device__ inline void narrow_primitive(const DevModel &m, EnvS &S, float *scratch, int lane) {
    for (int base = 0; base < S.ncand_p; base += 32) {
        int k = base + lane;
        PrimOut o;
        o.n = 0;
        int g1 = 0, g2 = 0;
        if (k < S.ncand_p) {
            int pk = S.cand_p[k], ty = (pk >> 16) & 0xff;
            g1 = pk & 0xff; g2 = (pk >> 8) & 0xff;
            Shape A = load_shape(m, S, g1), B = load_shape(m, S, g2);
            if (ty == AV_PAIR_SS) collide_sphere_sphere(A, B, o);
            else if (ty == AV_PAIR_SB) collide_sphere_box(A, B, o);
            else if (ty == AV_PAIR_BS) { collide_sphere_box(B, A, o); o.nrm = -o.nrm; }
            else if (!obb_separated(A.size, A, B.size, B)) collide_box_box(A, B, o);
        }
        // exclusive scan of counts
        int incl = o.n;
        for (int off = 1; off < 32; off <<= 1) {
            int v = __shfl_up_sync(AV_FULL, incl, off);
            if (lane >= off) incl += v;
        }
        int total = __shfl_sync(AV_FULL, incl, 31), start = S.ncon + incl - o.n;
        __syncwarp();
        for (int c = 0; c < o.n; c++) {
            if (start + c < AV_NCON) add_contact(m, S, scratch, start + c, g1, g2, o.dist[c], o.pos[c], o.nrm);
            else S.status |= 2;
        }
        if (lane == 0) S.ncon = min(AV_NCON, S.ncon + total);
        __syncwarp();
    }
    if (lane == 0) S.ncand_p = 0;
    __syncwarp();
}
// convex candidates (mesh hulls, cylinders): oriented-box rejection, one candidate per lane; survivors are appended to
// S.keep_c in candidate order; drains S.cand_c
__device__ inline void filter_convex(const DevModel &m, EnvS &S, int lane) {
    for (int base = 0; base < S.ncand_c; base += 32) {
        int k = base + lane, keep = 0, pk = 0;
        if (k < S.ncand_c) {
            pk = S.cand_c[k];
            int g1 = pk & 0xff, g2 = (pk >> 8) & 0xff;
            Shape A = load_shape(m, S, g1), B = load_shape(m, S, g2);
            keep = !obb_separated(ld3(m.geom_aabb + 3 * g1), A, ld3(m.geom_aabb + 3 * g2), B);
        }
        unsigned mk = __ballot_sync(AV_FULL, keep);
        int n0 = S.nkeep;
        __syncwarp();
        if (keep) {
            int s = n0 + __popc(mk & ((1u << lane) - 1u));
            if (s < AV_NKEEP) S.keep_c[s] = pk; else S.status |= 2;
        }
        if (lane == 0) S.nkeep = min(AV_NKEEP, n0 + __popc(mk));
        __syncwarp();
    }
    if (lane == 0) S.ncand_c = 0;
    __syncwarp();
}

// Collision, phase A (own environment): broadphase over the pair table, primitive narrowphase, box filter of the convex
// candidates.  The pair table lists all primitive pairs before all convex pairs (model compiler) and the primitive list
// is drained before anything convex is emitted, so contacts come out in pair-table order -- the order of the oracle's
// single loop.  The candidate lists are drained whenever they reach a warp's worth, so they cannot overflow.
__device__ AV_STAGE void stage_collision_a(const DevModel &m, EnvS &S, float *scratch, int lane, Prof &pf) {
    if (lane == 0) { S.ncon = 0; S.ncand_p = 0; S.ncand_c = 0; S.nkeep = 0; }
    __syncwarp();
    // broadphase: bounding spheres + world AABBs, pair list strided over lanes, warp-aggregated append
    for (int base = 0; base < m.npair; base += 32) {
        int p = base + lane, hit = 0, pk = 0;
        if (p < m.npair) {
            pk = m.pair_geom[p];
            int g1 = pk & 0xff, g2 = (pk >> 8) & 0xff;
            V3 d = ld3(S.gpos + 3 * g2) - ld3(S.gpos + 3 * g1);
            V3 h = ld3(S.gaabb + 3 * g1) + ld3(S.gaabb + 3 * g2);
            float rs = m.pair_rsum[p];
            hit = dot(d, d) <= rs * rs && fabsf(d.x) <= h.x && fabsf(d.y) <= h.y && fabsf(d.z) <= h.z;
        }
        int isconv = ((pk >> 16) & 0xff) == AV_PAIR_CONVEX;
        unsigned mp = __ballot_sync(AV_FULL, hit && !isconv), mc = __ballot_sync(AV_FULL, hit && isconv);
        if (!(mp | mc)) continue;
        int np = S.ncand_p, nc = S.ncand_c;
        __syncwarp();
        if (hit) {
            unsigned below = (1u << lane) - 1u;
            if (!isconv) S.cand_p[np + __popc(mp & below)] = pk;
            else S.cand_c[nc + __popc(mc & below)] = pk;
        }
        if (lane == 0) { S.ncand_p = np + __popc(mp); S.ncand_c = nc + __popc(mc); }
        __syncwarp();
        if (S.ncand_p >= 32) narrow_primitive(m, S, scratch, lane);
        if (S.ncand_c >= 32) filter_convex(m, S, lane);
    }
    pf.mark(PF_BROAD, lane);
    narrow_primitive(m, S, scratch, lane);
    filter_convex(m, S, lane);
    pf.mark(PF_PRIM, lane);
}

// Collision, phase B: pooled MPR runs.  An item is one run of one kept convex pair of environment S -- which may belong to
// ANOTHER warp of the block: the lockstep step kernel pools the runs of all its environments and lets every warp pull
// items, because the number of hull pairs per environment varies from 0 to 15 and the stage barrier would otherwise wait
// for the unluckiest.  Two rounds: the unperturbed runs (run = 0), then -- for the pairs that touch and take multiccd --
// the four perturbed runs (run = 1..4), which need the unperturbed contact point and normal.
// Result slots in the environment's scratch, per pair: hit | normal 3 | 5 x dist (1e30 = miss) | 5 x pos.
__device__ AV_STAGE void collide_item(const DevModel &m, const EnvS &S, float *scratch, int k, int run, int lane) {
    int pk = S.keep_c[k], g1 = pk & 0xff, g2 = (pk >> 8) & 0xff;
    Shape A = load_shape(m, S, g1), B = load_shape(m, S, g2);
    float *tmp = scratch + AV_SCR_TMP + k * AV_CTMP;
    V3 pos0 = v3(0, 0, 0), nrm0 = v3(0, 0, 1), pos, nrm;
    if (run > 0) { pos0 = ld3(tmp + 9); nrm0 = ld3(tmp + 1); }
    float dist;
    bool hit = collide_convex_run(A, B, run - 1, pos0, nrm0, lane, dist, pos, nrm);
    if (lane == 0) {
        if (run == 0) { tmp[0] = hit ? 1.f : 0.f; st3(tmp + 1, nrm); }
        tmp[4 + run] = hit ? dist : 1e30f;
        st3(tmp + 9 + 3 * run, pos);
    }
    __syncwarp();
}
// after the unperturbed round (own environment): the pairs that go on to the perturbed round -> S.cand_c[0..ncand_c)
__device__ inline void collide_select(const DevModel &m, EnvS &S, const float *scratch, int lane, bool multiccd) {
    for (int base = 0; base < S.nkeep; base += 32) {
        int k = base + lane, go = 0;
        if (k < S.nkeep && multiccd) {
            int pk = S.keep_c[k];
            go = scratch[AV_SCR_TMP + k * AV_CTMP] > 0.5f && m.geom_type[pk & 0xff] != AV_GEOM_SPHERE &&
                 m.geom_type[(pk >> 8) & 0xff] != AV_GEOM_SPHERE;
        }
        unsigned mk = __ballot_sync(AV_FULL, go);
        int n0 = S.ncand_c;
        __syncwarp();
        if (go) S.cand_c[n0 + __popc(mk & ((1u << lane) - 1u))] = k;
        if (lane == 0) S.ncand_c = n0 + __popc(mk);
        __syncwarp();
    }
}

// Collision, phase C (own environment): contacts from the pooled results, in candidate order; perturbed points that
// coincide with an earlier point of the pair (< 0.1 mm) are dropped, like collide_convex does
__device__ AV_STAGE void stage_collision_c(const DevModel &m, EnvS &S, float *scratch, int lane, bool multiccd, Prof &pf) {
    for (int k = 0; k < S.nkeep; k++) {
        const float *tmp = scratch + AV_SCR_TMP + k * AV_CTMP;
        if (!(tmp[0] > 0.5f)) continue;
        int pk = S.keep_c[k], g1 = pk & 0xff, g2 = (pk >> 8) & 0xff;
        bool mc = multiccd && m.geom_type[g1] != AV_GEOM_SPHERE && m.geom_type[g2] != AV_GEOM_SPHERE;
        V3 acc_pos[5];
        float acc_dist[5];
        int n = 1;
        acc_pos[0] = ld3(tmp + 9); acc_dist[0] = tmp[4];
#pragma unroll
        for (int r = 1; r <= 4; r++) {
            if (!mc || !(tmp[4 + r] < 1e29f)) continue;
            V3 ps = ld3(tmp + 9 + 3 * r);
            bool dup = false;
#pragma unroll
            for (int c = 0; c < 4; c++) dup = dup || (c < n && norm(ps - acc_pos[c]) < 1e-4f);
            if (dup) continue;
#pragma unroll
            for (int c = 1; c < 5; c++)
                if (c == n) { acc_pos[c] = ps; acc_dist[c] = tmp[4 + r]; }
            n++;
        }
        int n0 = S.ncon;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 5; c++)      // one lane writes each contact; unrolled so the small arrays stay in registers
            if (c < n && lane == c) {
                if (n0 + c < AV_NCON) add_contact(m, S, scratch, n0 + c, g1, g2, acc_dist[c], acc_pos[c], ld3(tmp + 1));
                else S.status |= 2;
            }
        if (lane == 0) S.ncon = min(AV_NCON, n0 + n);
        __syncwarp();
    }
    pf.mark(PF_CONVEX, lane);
}

// the phases back to back for one warp on its own (forward kernel, host emulation of single-warp blocks)
__device__ inline void stage_collision(const DevModel &m, EnvS &S, float *scratch, int lane, bool multiccd, Prof &pf) {
    stage_collision_a(m, S, scratch, lane, pf);
    for (int k = 0; k < S.nkeep; k++) collide_item(m, S, scratch, k, 0, lane);
    collide_select(m, S, scratch, lane, multiccd);
    for (int i = 0; i < 4 * S.ncand_c; i++) collide_item(m, S, scratch, S.cand_c[i >> 2], 1 + (i & 3), lane);
    stage_collision_c(m, S, scratch, lane, multiccd, pf);
}

// ------------------------------------------------------------------ K3b: velocity, bias, actuation, smooth acceleration
__device__ AV_STAGE void stage_smooth(const DevModel &m, EnvS &S, int lane) {
    if (lane < m.ntree) {
        int b0 = m.tree_bodyadr[lane], nb = m.tree_bodynum[lane];
        S6 zero = {v3(0, 0, 0), v3(0, 0, 0)};
        S6 grav = {v3(0, 0, 0), v3(-m.gravity[0], -m.gravity[1], -m.gravity[2])};
        for (int b = b0; b < b0 + nb; b++) {
            int p = m.body_parent[b];
            bool root = p < b0;
            S6 v = root ? zero : ld6(S.cvel + 6 * p);
            S6 a = root ? grav : ld6(S.crb + 12 * p);
            int d0 = m.body_dofadr[b], nd = m.body_dofnum[b];
            if (nd == 6) {
                for (int k = 0; k < 3; k++) {
                    st6(S.cdofdot + 6 * (d0 + k), zero);
                    v = v + ld6(S.cdof + 6 * (d0 + k)) * S.qvel[d0 + k];
                }
                for (int k = 3; k < 6; k++) {
                    S6 cd = cross_motion(v, ld6(S.cdof + 6 * (d0 + k)));
                    st6(S.cdofdot + 6 * (d0 + k), cd);
                    a = a + cd * S.qvel[d0 + k];
                }
                for (int k = 3; k < 6; k++) v = v + ld6(S.cdof + 6 * (d0 + k)) * S.qvel[d0 + k];
            } else {
                for (int k = 0; k < nd; k++) {
                    S6 c = ld6(S.cdof + 6 * (d0 + k));
                    S6 cd = cross_motion(v, c);
                    st6(S.cdofdot + 6 * (d0 + k), cd);
                    a = a + cd * S.qvel[d0 + k];
                    v = v + c * S.qvel[d0 + k];
                }
            }
            st6(S.cvel + 6 * b, v);
            st6(S.crb + 12 * b, a);
            const float *I = S.cinert + 10 * b;
            S6 f = inert_mul(I, a) + cross_force(v, inert_mul(I, v));
            st6(S.crb + 12 * b + 6, f);
        }
        for (int b = b0 + nb - 1; b > b0; b--) {
            int p = m.body_parent[b];
            st6(S.crb + 12 * p + 6, ld6(S.crb + 12 * p + 6) + ld6(S.crb + 12 * b + 6));
        }
    }
    __syncwarp();
    for (int i = lane; i < m.nv; i += 32) {
        float bias = dot6(ld6(S.cdof + 6 * i), ld6(S.crb + 12 * m.dof_body[i] + 6));
        S.qfrc_bias[i] = bias;
        S.qfrc_smooth[i] = -m.dof_damping[i] * S.qvel[i] - bias;
    }
    __syncwarp();
    if (lane < m.nu) {  // position servos; every actuator drives its own dof
        int u = lane, i = m.act_dof[u];
        float c = fminf(fmaxf(S.ctrl[u], m.act_ctrl_lo[u]), m.act_ctrl_hi[u]);
        float f = m.act_kp[u] * (c - S.qpos[m.act_qadr[u]]) - m.act_kv[u] * S.qvel[i];
        if (m.dof_frc_limited[i]) f = fminf(fmaxf(f, m.dof_frc_lo[i]), m.dof_frc_hi[i]);
        S.qfrc_smooth[i] += f;
    }
    __syncwarp();
    for (int i = lane; i < m.nv; i += 32) {
        int t = m.dof_tree[i], d0 = m.tree_dofadr[t], nt = m.tree_dofnum[t];
        const float *row = S.Minv + t * AV_TD * AV_TD + (i - d0) * AV_TD;
        float s = 0.f;
        for (int k = 0; k < nt; k++) s += row[k] * S.qfrc_smooth[d0 + k];
        S.qacc_smooth[i] = s;
    }
    __syncwarp();
}

// ------------------------------------------------------------------ K5: constraint rows
__device__ __forceinline__ float impedance(const float *solimp, float x_abs) {
    float d0 = solimp[0], dw = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
    if (width < AV_MINVAL) return 0.5f * (d0 + dw);
    float x = x_abs / width;
    if (x >= 1.f) return dw;
    if (x <= 0.f) return d0;
    float y;
    if (power == 1.f) y = x;
    else if (power == 2.f) {   // the default solimp (every constraint of the AV-ALOHA models): no powf
        float a = x <= mid ? x / mid : (1.f - x) / (1.f - mid);
        y = x <= mid ? a * a * mid : 1.f - a * a * (1.f - mid);
    } else if (x <= mid) y = powf(x / mid, power) * mid;
    else y = 1.f - powf((1.f - x) / (1.f - mid), power) * (1.f - mid);
    return d0 + y * (dw - d0);
}
__device__ __forceinline__ void kbi(const DevModel &m, const float *solref, const float *solimp, float pos, float &K,
                                    float &B, float &imp) {
    float tc = fmaxf(solref[0], 2.f * m.timestep), dr = solref[1], dmax = solimp[1];
    imp = impedance(solimp, fabsf(pos));
    B = 2.f / (dmax * tc);
    K = 1.f / (dmax * dmax * tc * tc * dr * dr);
}

// one scalar row (equality / friction loss / joint limit), written by the lane that found it
__device__ inline void emit_scalar_row(const DevModel &m, EnvS &S, const FCache &fc, int key, int r, int d1, int d2, float c1, float c2, float pos,
                                       float lo, float hi, const float *solref, const float *solimp, float invw) {
    if (r >= AV_NSC) { S.status |= 4; return; }
    float K, B, imp;
    kbi(m, solref, solimp, pos, K, B, imp);
    int t = m.dof_tree[d1], d0 = m.tree_dofadr[t], nt = m.tree_dofnum[t];
    const float *Mi = S.Minv + t * AV_TD * AV_TD;
    float vel = c1 * S.qvel[d1] + (d2 >= 0 ? c2 * S.qvel[d2] : 0.f);
    float aref = -B * vel - K * imp * pos;
    for (int k = 0; k < AV_TD; k++) {
        float v = 0.f;
        if (k < nt) v = c1 * Mi[(d1 - d0) * AV_TD + k] + (d2 >= 0 ? c2 * Mi[(d2 - d0) * AV_TD + k] : 0.f);
        S.sc_MJ[r * AV_TD + k] = v;
    }
    float A = c1 * S.sc_MJ[r * AV_TD + (d1 - d0)] + (d2 >= 0 ? c2 * S.sc_MJ[r * AV_TD + (d2 - d0)] : 0.f);
    float R = fmaxf(AV_MINVAL, (1.f - imp) / imp * invw);
    S.sc_dof1[r] = d1; S.sc_dof2[r] = d2; S.sc_tree[r] = t; S.sc_c1[r] = c1; S.sc_c2[r] = c2;
    S.sc_aref[r] = aref;
    S.sc_b[r] = c1 * S.qacc_smooth[d1] + (d2 >= 0 ? c2 * S.qacc_smooth[d2] : 0.f) - aref;
    S.sc_R[r] = R; S.sc_A[r] = A; S.sc_lo[r] = lo; S.sc_hi[r] = hi;
    // warm start: primal -> dual map, clamped
    float jw = c1 * S.warm[d1] + (d2 >= 0 ? c2 * S.warm[d2] : 0.f);
    float f0 = -(jw - aref) / R;
    if (fc.mode == 2) {   // force of the same constraint in the previous solve (0 when it is new)
        f0 = 0.f;
        int ns = fc.n[1];
        for (int k = 0; k < ns; k++)
            if (fc.key[AV_NCON + k] == key) { f0 = fc.val[AV_NCON * 6 + k]; break; }
    }
    S.sc_key[r] = key;
    S.sc_f[r] = fminf(fmaxf(f0, lo), hi);
}

// rows in MuJoCo's order [equality | friction loss | violated joint limits]; row indices via ballots
__device__ AV_STAGE void stage_rows_scalar(const DevModel &m, EnvS &S, int lane, const FCache &fc) {
    const float BIG = 3.0e38f;
    if (lane < m.neq) {
        int e = lane;
        const float *c = m.eq_polycoef + 5 * e;
        float x = S.qpos[m.eq_qadr2[e]] - m.qpos0[m.eq_qadr2[e]];
        float poly = c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * c[4])));
        float dpoly = c[1] + x * (2 * c[2] + x * (3 * c[3] + x * 4 * c[4]));
        float pos = S.qpos[m.eq_qadr1[e]] - m.qpos0[m.eq_qadr1[e]] - poly;
        emit_scalar_row(m, S, fc, (0 << 24) | e, e, m.eq_dof1[e], m.eq_dof2[e], 1.f, -dpoly, pos, -BIG, BIG, m.eq_solref + 2 * e,
                        m.eq_solimp + 5 * e, m.eq_invweight0[e]);
    }
    int nrow = m.neq;
    for (int base = 0; base < m.nv; base += 32) {
        int i = base + lane;
        int has = i < m.nv && m.dof_frictionloss[i] > 0.f;
        unsigned mk = __ballot_sync(AV_FULL, has);
        if (has) {
            float fl = m.dof_frictionloss[i];
            emit_scalar_row(m, S, fc, (1 << 24) | i, nrow + __popc(mk & ((1u << lane) - 1u)), i, -1, 1.f, 0.f, 0.f, -fl, fl,
                            m.dof_solref + 2 * i, m.dof_solimp + 5 * i, m.dof_invweight0[i]);
        }
        nrow += __popc(mk);
    }
    for (int base = 0; base < 2 * m.njnt; base += 32) {
        int w = base + lane, j = w >> 1, side = w & 1;
        int has = 0;
        float dist = 0.f;
        if (w < 2 * m.njnt && m.jnt_limited[j] && m.jnt_type[j] != AV_JNT_FREE) {
            float q = S.qpos[m.jnt_qposadr[j]];
            dist = side ? m.jnt_range[2 * j + 1] - q : q - m.jnt_range[2 * j];
            has = dist < 0.f;
        }
        unsigned mk = __ballot_sync(AV_FULL, has);
        if (has) {
            int d1 = m.jnt_dofadr[j];
            emit_scalar_row(m, S, fc, (2 << 24) | j, nrow + __popc(mk & ((1u << lane) - 1u)), d1, -1, side ? -1.f : 1.f, 0.f, dist, 0.f, BIG,
                            m.jnt_solref + 2 * j, m.jnt_solimp + 5 * j, m.dof_invweight0[d1]);
        }
        nrow += __popc(mk);
    }
    if (lane == 0) S.nsc = nrow < AV_NSC ? nrow : AV_NSC;
    __syncwarp();
}

#include "avsim_solve.cuh"
#include "avsim_newton.cuh"

// ------------------------------------------------------------------ K7: Euler with implicit joint damping
__device__ AV_STAGE void stage_integrate(const DevModel &m, EnvS &S, int lane) {
    float h = m.timestep;
    // total generalized force = smooth + constraint = smooth + M * acc   (acc = M^-1 J^T f)
    float tot = 0.f;
    int i = lane, i2 = lane + 32;
    float tot2 = 0.f;
    for (int pass = 0; pass < 2; pass++) {
        int d = pass ? i2 : i;
        if (d < m.nv) {
            int t = m.dof_tree[d], d0 = m.tree_dofadr[t], nt = m.tree_dofnum[t];
            const float *Mb = S.M + t * AV_MTRI;
            float s = S.qfrc_smooth[d];
            for (int k = 0; k < nt; k++) s += Mb[av_mtri(d - d0, k)] * S.acc[d0 + k];
            if (pass) tot2 = s; else tot = s;
        }
    }
    __syncwarp();
    if (i < m.nv) { S.warm[i] = S.qacc_smooth[i] + S.acc[i]; S.qfrc_bias[i] = tot; }   // qfrc_bias reused as rhs
    if (i2 < m.nv) { S.warm[i2] = S.qacc_smooth[i2] + S.acc[i2]; S.qfrc_bias[i2] = tot2; }
    __syncwarp();
    if (lane < m.ntree) {
        int nt = m.tree_dofnum[lane], d0 = m.tree_dofadr[lane];
        float *Lb = S.L + lane * AV_TD * AV_TD;
        chol_block(S.M + lane * AV_MTRI, Lb, nt, m.dof_damping + d0, h);
        float x[AV_TD];
        for (int k = 0; k < nt; k++) x[k] = S.qfrc_bias[d0 + k];
        chol_block_solve(Lb, nt, x);
        for (int k = 0; k < nt; k++) S.qvel[d0 + k] += h * x[k];
    }
    __syncwarp();
    for (int j = lane; j < m.njnt; j += 32) {
        int qa = m.jnt_qposadr[j], da = m.jnt_dofadr[j];
        if (m.jnt_type[j] == AV_JNT_FREE) {
            for (int k = 0; k < 3; k++) S.qpos[qa + k] += h * S.qvel[da + k];
            V3 w = ld3(S.qvel + da + 3);
            float n = norm(w), ang = n * h;
            V3 ax = n > AV_MINVAL ? w * (1.f / n) : w;
            float s = sinf(0.5f * ang);
            Q4 rq = {cosf(0.5f * ang), s * ax.x, s * ax.y, s * ax.z};
            stq(S.qpos + qa + 3, qnormalize(qmul(ldq(S.qpos + qa + 3), rq)));
        } else
            S.qpos[qa] += h * S.qvel[da];
    }
    __syncwarp();
}

// ------------------------------------------------------------------ K8: reward from contact classes
enum { CLS_LEFT = 1, CLS_RIGHT = 2, CLS_TABLE = 4, CLS_A = 8, CLS_B = 16, CLS_PIN_A = 32, CLS_PIN_B = 64, CLS_C = 128 };
__device__ __forceinline__ int pair_hit(int c1, int c2, int ma, int mb) {
    return ((c1 & ma) && (c2 & mb)) || ((c2 & ma) && (c1 & mb));
}
__device__ AV_STAGE int stage_reward(const DevModel &m, const EnvS &S, int lane, int &latch) {
    int flags = 0;  // bit0 tl, 1 tr, 2 a_table, 3 b_table, 4 a_b, 5 pins, 6 a_pinb
    for (int c = lane; c < S.ncon; c += 32) {
        int info = S.c_info[c], c1 = m.geom_class[info & 0xff], c2 = m.geom_class[(info >> 8) & 0xff];
        int t = m.task_id;
        int handA = (t == 0 || t == 3) ? CLS_B : CLS_A;   // which object class the *left* hand must touch
        flags |= pair_hit(c1, c2, CLS_A, CLS_RIGHT) << 1;
        flags |= pair_hit(c1, c2, handA, CLS_LEFT);
        flags |= pair_hit(c1, c2, CLS_TABLE, CLS_A) << 2;
        flags |= pair_hit(c1, c2, CLS_TABLE, CLS_B) << 3;
        flags |= pair_hit(c1, c2, CLS_A, CLS_B) << 4;
        int pins = (t == 0) ? pair_hit(c1, c2, CLS_A, CLS_PIN_A)
                 : (t == 3) ? pair_hit(c1, c2, CLS_C, CLS_PIN_A) : pair_hit(c1, c2, CLS_PIN_A, CLS_PIN_B);
        flags |= pins << 5;
        flags |= pair_hit(c1, c2, CLS_A, CLS_PIN_B) << 6;
    }
    for (int o = 16; o > 0; o >>= 1) flags |= __shfl_xor_sync(AV_FULL, flags, o);
    int tl = flags & 1, tr = (flags >> 1) & 1, at = (flags >> 2) & 1, bt = (flags >> 3) & 1, ab = (flags >> 4) & 1,
        pins = (flags >> 5) & 1, apb = (flags >> 6) & 1, r = 0;
    switch (m.task_id) {
    case 0:
        if (tl && tr) r = 1;
        if (tl && tr && !at && !bt) r = 2;
        if (ab && !at && !bt) r = 3;
        if (pins) r = 4;
        break;
    case 2:
        if (pins) latch = 1;
        if (tr) r = 1;
        if (tr && !at) r = 2;
        if (ab && !at) r = 3;
        if (latch) r = 4;
        if (tl && !tr && !at && !apb && latch) r = 5;
        break;
    case 3:
        if (tl && tr) r = 1;
        if (tl && tr && !at && !bt) r = 2;
        if (pins) r = 3;
        break;
    default:  // slot insertion (1) and hook package (4) share the staging
        if (tl && tr) r = 1;
        if (tl && tr && !at) r = 2;
        if (ab && !at) r = 3;
        if (pins) r = 4;
        break;
    }
    return r;
}
