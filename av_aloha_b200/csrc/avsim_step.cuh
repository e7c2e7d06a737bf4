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

struct EnvS {
    float qpos[AV_NQ], qvel[AV_NVP], ctrl[24], warm[AV_NVP];
    float xpos[AV_NB * 3], xquat[AV_NB * 4], xmat[AV_NB * 9], xipos[AV_NB * 3];
    float torig[AV_NTREE * 3];
    float cdof[AV_NV * 6], cdofdot[AV_NV * 6];
    float cinert[AV_NB * 10];
    float crb[AV_NB * 12];  // composite inertias (10/body) in stage_inertia; cacc|cfrc (6+6/body) in stage_smooth
    float cvel[AV_NB * 6];
    float M[AV_MBLK], L[AV_MBLK], Minv[AV_MBLK];
    float qfrc_smooth[AV_NVP], qacc_smooth[AV_NVP], acc[AV_NVP], qfrc_bias[AV_NVP];
    float gpos[AV_NG * 3], gaabb[AV_NG * 3];  // world centre + world-axis half extents of every geom
    // contacts
    float c_pos[AV_NCON * 3], c_frame[AV_NCON * 9], c_dist[AV_NCON], c_mu[AV_NCON * 3], c_b[AV_NCON * 6],
        c_f[AV_NCON * 6], c_Rn[AV_NCON], c_aref[AV_NCON * 6];
    int c_info[AV_NCON];  // geom1 | geom2 << 8 | dim << 16 | excluded << 20
    int c_tree[AV_NCON];  // tree1 | tree2 << 8 (0xff: none)
    // scalar rows
    int sc_dof1[AV_NSC], sc_dof2[AV_NSC], sc_tree[AV_NSC];
    float sc_c1[AV_NSC], sc_c2[AV_NSC], sc_b[AV_NSC], sc_R[AV_NSC], sc_f[AV_NSC], sc_lo[AV_NSC], sc_hi[AV_NSC],
        sc_A[AV_NSC], sc_aref[AV_NSC], sc_MJ[AV_NSC * AV_TD];
    float stage[2 * 6 * AV_JW];  // J | MinvJT of the contact being assembled
    int cand_p[AV_NCAND], cand_c[AV_NCAND];
    int ncon, nsc, ncand_p, ncand_c, status;
};

__device__ __forceinline__ int body_mask_has(const DevModel &m, int body, int dof) {
    // is `dof` on the path from the tree root to `body`?
    for (int i = m.body_lastdof[body]; i >= 0; i = m.dof_parent[i])
        if (i == dof) return 1;
    return 0;
}

// ------------------------------------------------------------------ K2: kinematics, lane = kinematic tree
__device__ inline void stage_kinematics(const DevModel &m, EnvS &S, int lane) {
    if (lane < m.ntree) {
        int b0 = m.tree_bodyadr[lane], nb = m.tree_bodynum[lane];
        V3 org = v3(0, 0, 0);
        for (int b = b0; b < b0 + nb; b++) {
            int p = m.body_parent[b];
            M3 Rp = ldm3(S.xmat + 9 * p);
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
            st3(S.xpos + 3 * b, pos); stq(S.xquat + 4 * b, quat); stm3(S.xmat + 9 * b, R);
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
    for (int g = lane; g < m.ngeom; g += 32)
        if (!m.geom_static[g]) {
            int b = m.geom_body[g];
            M3 Rb = ldm3(S.xmat + 9 * b);
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
// Cholesky of the nt x nt block at A (stride AV_TD) into Lb (lower); returns false on a non-positive pivot
__device__ inline bool chol_block(const float *A, float *Lb, int nt, const float *diag_add, float h) {
    bool ok = true;
    for (int i = 0; i < nt; i++)
        for (int j = 0; j <= i; j++) {
            float s = A[i * AV_TD + j];
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

__device__ inline void stage_inertia(const DevModel &m, EnvS &S, int lane) {
    for (int i = lane; i < AV_MBLK; i += 32) S.M[i] = 0.f;
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
        float *Mb = S.M + t * AV_TD * AV_TD;
        for (int j = i; j >= 0; j = m.dof_parent[j]) {
            float v = dot6(ld6(S.cdof + 6 * j), f);
            if (j == i) v += m.dof_armature[i];
            Mb[(i - d0) * AV_TD + (j - d0)] = v;
            Mb[(j - d0) * AV_TD + (i - d0)] = v;
        }
    }
    __syncwarp();
    if (lane < m.ntree) {
        int nt = m.tree_dofnum[lane];
        if (!chol_block(S.M + lane * AV_TD * AV_TD, S.L + lane * AV_TD * AV_TD, nt, nullptr, 0.f)) S.status |= 1;
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
        s.mat = mul(ldm3(S.xmat + 9 * b), ldm3(m.geom_mat + 9 * g));
    }
    if (s.type == AV_GEOM_MESH) {
        int h = m.geom_hull[g];
        s.vert = m.hull_vert + m.hull_adr[h];
        s.nvert = m.hull_num[h];
    }
    return s;
}

__device__ inline void add_contact(const DevModel &m, EnvS &S, int slot, int g1, int g2, float dist, V3 pos, V3 nrm) {
    V3 t1, t2;
    make_frame(nrm, t1, t2);
    st3(S.c_pos + 3 * slot, pos);
    st3(S.c_frame + 9 * slot, nrm); st3(S.c_frame + 9 * slot + 3, t1); st3(S.c_frame + 9 * slot + 6, t2);
    S.c_dist[slot] = dist;
    int dim = max(m.geom_condim[g1], m.geom_condim[g2]);
    for (int k = 0; k < 3; k++) S.c_mu[3 * slot + k] = fmaxf(m.geom_friction[3 * g1 + k], m.geom_friction[3 * g2 + k]);
    float margin = fmaxf(m.geom_margin[g1], m.geom_margin[g2]), gap = fmaxf(m.geom_gap[g1], m.geom_gap[g2]);
    int excluded = !(dist < margin - gap);
    S.c_info[slot] = g1 | (g2 << 8) | (dim << 16) | (excluded << 20);
}

__device__ inline void stage_collision(const DevModel &m, EnvS &S, int lane, bool multiccd) {
    if (lane == 0) { S.ncon = 0; S.ncand_p = 0; S.ncand_c = 0; }
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
        int np = S.ncand_p, nc = S.ncand_c;
        __syncwarp();
        if (hit) {
            unsigned below = (1u << lane) - 1u;
            if (!isconv) { int s = np + __popc(mp & below); if (s < AV_NCAND) S.cand_p[s] = pk; else S.status |= 2; }
            else { int s = nc + __popc(mc & below); if (s < AV_NCAND) S.cand_c[s] = pk; else S.status |= 2; }
        }
        if (lane == 0) { S.ncand_p = min(AV_NCAND, np + __popc(mp)); S.ncand_c = min(AV_NCAND, nc + __popc(mc)); }
        __syncwarp();
    }
    // primitive pairs: one candidate per lane
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
            if (start + c < AV_NCON) add_contact(m, S, start + c, g1, g2, o.dist[c], o.pos[c], o.nrm);
            else S.status |= 2;
        }
        if (lane == 0) S.ncon = min(AV_NCON, S.ncon + total);
        __syncwarp();
    }
    // convex pairs: oriented-box rejection per lane, then warp-cooperative MPR one pair at a time
    for (int base = 0; base < S.ncand_c; base += 32) {
        int k = base + lane, keep = 0;
        if (k < S.ncand_c) {
            int pk = S.cand_c[k], g1 = pk & 0xff, g2 = (pk >> 8) & 0xff;
            Shape A = load_shape(m, S, g1), B = load_shape(m, S, g2);
            keep = !obb_separated(ld3(m.geom_aabb + 3 * g1), A, ld3(m.geom_aabb + 3 * g2), B);
        }
        unsigned mk = __ballot_sync(AV_FULL, keep);
        while (mk) {
            int src = __ffs(mk) - 1;
            mk &= mk - 1;
            int pk = S.cand_c[base + src], g1 = pk & 0xff, g2 = (pk >> 8) & 0xff;
            Shape A = load_shape(m, S, g1), B = load_shape(m, S, g2);
            PrimOut o;
            bool mc = multiccd && A.type != AV_GEOM_SPHERE && B.type != AV_GEOM_SPHERE;
            collide_convex(A, B, mc, lane, o);
            int n0 = S.ncon;
            __syncwarp();
            if (lane < o.n) {
                if (n0 + lane < AV_NCON) add_contact(m, S, n0 + lane, g1, g2, o.dist[lane], o.pos[lane], o.nrm);
                else S.status |= 2;
            }
            if (lane == 0) S.ncon = min(AV_NCON, n0 + o.n);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ K3b: velocity, bias, actuation, smooth acceleration
__device__ inline void stage_smooth(const DevModel &m, EnvS &S, int lane) {
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
    else if (x <= mid) y = powf(x / mid, power) * mid;
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
__device__ inline void emit_scalar_row(const DevModel &m, EnvS &S, int r, int d1, int d2, float c1, float c2, float pos,
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
    S.sc_f[r] = fminf(fmaxf(-(jw - aref) / R, lo), hi);
}

// rows in MuJoCo's order [equality | friction loss | violated joint limits]; row indices via ballots
__device__ inline void stage_rows_scalar(const DevModel &m, EnvS &S, int lane) {
    const float BIG = 3.0e38f;
    if (lane < m.neq) {
        int e = lane;
        const float *c = m.eq_polycoef + 5 * e;
        float x = S.qpos[m.eq_qadr2[e]] - m.qpos0[m.eq_qadr2[e]];
        float poly = c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * c[4])));
        float dpoly = c[1] + x * (2 * c[2] + x * (3 * c[3] + x * 4 * c[4]));
        float pos = S.qpos[m.eq_qadr1[e]] - m.qpos0[m.eq_qadr1[e]] - poly;
        emit_scalar_row(m, S, e, m.eq_dof1[e], m.eq_dof2[e], 1.f, -dpoly, pos, -BIG, BIG, m.eq_solref + 2 * e,
                        m.eq_solimp + 5 * e, m.eq_invweight0[e]);
    }
    int nrow = m.neq;
    for (int base = 0; base < m.nv; base += 32) {
        int i = base + lane;
        int has = i < m.nv && m.dof_frictionloss[i] > 0.f;
        unsigned mk = __ballot_sync(AV_FULL, has);
        if (has) {
            float fl = m.dof_frictionloss[i];
            emit_scalar_row(m, S, nrow + __popc(mk & ((1u << lane) - 1u)), i, -1, 1.f, 0.f, 0.f, -fl, fl,
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
            emit_scalar_row(m, S, nrow + __popc(mk & ((1u << lane) - 1u)), d1, -1, side ? -1.f : 1.f, 0.f, dist, 0.f, BIG,
                            m.jnt_solref + 2 * j, m.jnt_solimp + 5 * j, m.dof_invweight0[d1]);
        }
        nrow += __popc(mk);
    }
    if (lane == 0) S.nsc = nrow < AV_NSC ? nrow : AV_NSC;
    __syncwarp();
}

// packed lower-triangular index
__device__ __forceinline__ int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// J . x for the 3 rows this lane holds, reduced over the 16 columns of its half; lane 0 / lane 16 hold the sums
__device__ __forceinline__ void block_dot(const float j0, const float j1, const float j2, float x, float &r0, float &r1,
                                          float &r2) {
    r0 = j0 * x; r1 = j1 * x; r2 = j2 * x;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        r0 += __shfl_xor_sync(AV_FULL, r0, o);
        r1 += __shfl_xor_sync(AV_FULL, r1, o);
        r2 += __shfl_xor_sync(AV_FULL, r2, o);
    }
}

// contacts, one at a time, lane = (half = lane >> 4 -> rows 3*half..3*half+2, col = lane & 15)
__device__ inline void stage_rows_contact(const DevModel &m, EnvS &S, float *scratch, int lane) {
    int col = lane & 15, half = lane >> 4;
    for (int c = 0; c < S.ncon; c++) {
        int info = S.c_info[c], g1 = info & 0xff, g2 = (info >> 8) & 0xff, dim = (info >> 16) & 0xf;
        if ((info >> 20) & 1) { if (lane == 0) S.c_tree[c] = 0xffff; continue; }
        int b1 = m.geom_body[g1], b2 = m.geom_body[g2];
        int t1 = m.body_tree[b1], t2 = m.body_tree[b2];
        if (t1 < 0) { t1 = t2; t2 = -1; }   // keep the first slot occupied; signs are handled per body below
        if (t2 == t1) t2 = -1;
        int t = col < 8 ? t1 : t2, dl = col & 7, dof = -1;
        if (t >= 0 && dl < m.tree_dofnum[t]) dof = m.tree_dofadr[t] + dl;
        float j0 = 0.f, j1 = 0.f, j2 = 0.f;
        V3 p = ld3(S.c_pos + 3 * c);
        if (dof >= 0) {
            float sgn = 0.f;
            if (m.body_tree[b2] == t && body_mask_has(m, b2, dof)) sgn += 1.f;
            if (m.body_tree[b1] == t && body_mask_has(m, b1, dof)) sgn -= 1.f;
            if (sgn != 0.f) {
                S6 cd = ld6(S.cdof + 6 * dof);
                V3 v = half == 0 ? cd.l + cross(cd.a, p - ld3(S.torig + 3 * t)) : cd.a;
                v = v * sgn;
                j0 = dot(ld3(S.c_frame + 9 * c), v);
                j1 = dot(ld3(S.c_frame + 9 * c + 3), v);
                j2 = dot(ld3(S.c_frame + 9 * c + 6), v);
            }
        }
        if (half == 1 && dim < 6) { j0 = j1 = j2 = 0.f; }   // condim 3: no torsional / rolling rows
        float *Js = S.stage, *MJs = S.stage + 6 * AV_JW;
        Js[(3 * half + 0) * AV_JW + col] = j0; Js[(3 * half + 1) * AV_JW + col] = j1; Js[(3 * half + 2) * AV_JW + col] = j2;
        __syncwarp();
        // MinvJT rows: (J row) x (block inverse of the column's tree)
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;
        if (dof >= 0) {
            const float *Mi = S.Minv + t * AV_TD * AV_TD + dl;  // column dl (symmetric)
            int cb = col & 8, nt = m.tree_dofnum[t];
            for (int k = 0; k < nt; k++) {
                float mv = Mi[k * AV_TD];
                m0 += Js[(3 * half + 0) * AV_JW + cb + k] * mv;
                m1 += Js[(3 * half + 1) * AV_JW + cb + k] * mv;
                m2 += Js[(3 * half + 2) * AV_JW + cb + k] * mv;
            }
        }
        MJs[(3 * half + 0) * AV_JW + col] = m0; MJs[(3 * half + 1) * AV_JW + col] = m1; MJs[(3 * half + 2) * AV_JW + col] = m2;
        float *blk = scratch + c * AV_CBLK;
        blk[(3 * half + 0) * AV_JW + col] = j0; blk[(3 * half + 1) * AV_JW + col] = j1; blk[(3 * half + 2) * AV_JW + col] = j2;
        blk[6 * AV_JW + (3 * half + 0) * AV_JW + col] = m0; blk[6 * AV_JW + (3 * half + 1) * AV_JW + col] = m1;
        blk[6 * AV_JW + (3 * half + 2) * AV_JW + col] = m2;
        // velocities / smooth accelerations / warm-start accelerations along the rows
        float xv = dof >= 0 ? S.qvel[dof] : 0.f, xa = dof >= 0 ? S.qacc_smooth[dof] : 0.f, xw = dof >= 0 ? S.warm[dof] : 0.f;
        float v0, v1, v2, a0, a1, a2, w0, w1, w2;
        block_dot(j0, j1, j2, xv, v0, v1, v2);
        block_dot(j0, j1, j2, xa, a0, a1, a2);
        block_dot(j0, j1, j2, xw, w0, w1, w2);
        __syncwarp();
        // impedance, regularisation, reference acceleration
        float K, B, imp, solref[2], solimp[5];
        for (int k = 0; k < 2; k++) solref[k] = 0.5f * (m.geom_solref[2 * g1 + k] + m.geom_solref[2 * g2 + k]);
        for (int k = 0; k < 5; k++) solimp[k] = 0.5f * (m.geom_solimp[5 * g1 + k] + m.geom_solimp[5 * g2 + k]);
        float dist = S.c_dist[c];
        kbi(m, solref, solimp, dist, K, B, imp);
        float Rn = fmaxf(AV_MINVAL, (1.f - imp) / imp * (m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2]));
        float mu0 = S.c_mu[3 * c], mu1 = S.c_mu[3 * c + 1], mu2 = S.c_mu[3 * c + 2];
        float Rf = Rn / m.impratio;
        float R[6] = {Rn, Rf, Rf, Rf * mu0 * mu0 / (mu1 * mu1), Rf * mu0 * mu0 / (mu2 * mu2), Rf * mu0 * mu0 / (mu2 * mu2)};
        if (col == 0) {  // lanes 0 and 16 hold the reduced sums of their three rows
            float vv[3] = {v0, v1, v2}, aa[3] = {a0, a1, a2}, ww[3] = {w0, w1, w2};
            for (int k = 0; k < 3; k++) {
                int row = 3 * half + k;
                float aref = -B * vv[k] - (row == 0 ? K * imp * dist : 0.f);
                bool live = row < dim;
                S.c_aref[6 * c + row] = live ? aref : 0.f;
                S.c_b[6 * c + row] = live ? aa[k] - aref : 0.f;
                S.c_f[6 * c + row] = live ? -(ww[k] - aref) / R[row] : 0.f;
            }
        }
        if (lane == 0) { S.c_Rn[c] = Rn; S.c_tree[c] = (t1 & 0xff) | ((t2 & 0xff) << 8); }
        // AR = J MinvJT^T + R, packed lower triangle (21 entries), lane = entry
        if (lane < 21) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= lane) i++;
            int j = lane - i * (i + 1) / 2;
            float s = 0.f;
            for (int k = 0; k < AV_JW; k++) s += Js[i * AV_JW + k] * MJs[j * AV_JW + k];
            if (i == j) s += (i < dim) ? R[i] : 1.f;   // dead rows get a unit diagonal
            blk[12 * AV_JW + lane] = s;
        }
        __syncwarp();
    }
    // lane = contact: Cholesky factor of the regularised friction block (rows 1..dim-1), packed lower (15)
    for (int c = lane; c < S.ncon; c += 32) {
        int info = S.c_info[c], dim = (info >> 16) & 0xf;
        if ((info >> 20) & 1) continue;
        float *blk = scratch + c * AV_CBLK;
        const float *AR = blk + 12 * AV_JW;
        float Lc[15];
        int n = dim - 1;
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++) {
                float s = AR[tri(i + 1, j + 1)];
                for (int k = 0; k < j; k++) s -= Lc[tri(i, k)] * Lc[tri(j, k)];
                Lc[tri(i, j)] = (i == j) ? sqrtf(fmaxf(s, AV_MINVAL)) : s / Lc[tri(j, j)];
            }
        for (int k = 0; k < n * (n + 1) / 2; k++) blk[12 * AV_JW + 21 + k] = Lc[k];
        // project the warm-start force onto the cone
        float *f = S.c_f + 6 * c;
        if (f[0] <= 0.f) { for (int k = 0; k < 6; k++) f[k] = 0.f; }
        else {
            float mu[5] = {S.c_mu[3 * c], S.c_mu[3 * c], S.c_mu[3 * c + 1], S.c_mu[3 * c + 2], S.c_mu[3 * c + 2]};
            float s = 0.f;
            for (int k = 1; k < dim; k++) s += (f[k] / mu[k - 1]) * (f[k] / mu[k - 1]);
            if (s > f[0] * f[0]) { float sc = f[0] * rsqrtf(s); for (int k = 1; k < dim; k++) f[k] *= sc; }
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------ K6: block projected Gauss-Seidel on the dual + noslip
// solve Lc Lc^T x = -b (n <= 5, Lc packed lower)
__device__ __forceinline__ void tri_solve(const float *Lc, int n, const float *b, float *x) {
    for (int i = 0; i < n; i++) {
        float s = -b[i];
        for (int k = 0; k < i; k++) s -= Lc[tri(i, k)] * x[k];
        x[i] = s / Lc[tri(i, i)];
    }
    for (int i = n - 1; i >= 0; i--) {
        float s = x[i];
        for (int k = i + 1; k < n; k++) s -= Lc[tri(k, i)] * x[k];
        x[i] = s / Lc[tri(i, i)];
    }
}
__device__ inline void chol_small(const float A[5][5], int n, float lam, float *Lc) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            float s = A[i][j] + (i == j ? lam : 0.f);
            for (int k = 0; k < j; k++) s -= Lc[tri(i, k)] * Lc[tri(j, k)];
            Lc[tri(i, j)] = (i == j) ? sqrtf(fmaxf(s, AV_MINVAL)) : s / Lc[tri(j, j)];
        }
}
// minimise 0.5 y'Ay + y'b  s.t.  sum (y_i/mu_i)^2 <= r^2.  Lc0 (nullable) = Cholesky factor of A itself.
__device__ inline void qcqp(int n, const float Ain[5][5], const float *bin, const float *mu, float r, const float *Lc0,
                            float *y) {
    float Lc[15], z[5];
    if (Lc0) tri_solve(Lc0, n, bin, y);
    else { chol_small(Ain, n, 0.f, Lc); tri_solve(Lc, n, bin, y); }
    float zz = 0.f;
    for (int i = 0; i < n; i++) zz += (y[i] / mu[i]) * (y[i] / mu[i]);
    if (zz <= r * r) return;
    // scaled problem z = y / mu;  Newton on the multiplier of |z| = r
    float A[5][5], b[5], w[5], mz[5];
    for (int i = 0; i < n; i++) {
        b[i] = bin[i] * mu[i];
        for (int j = 0; j < n; j++) A[i][j] = Ain[i][j] * mu[i] * mu[j];
    }
    float lam = 0.f;
    for (int it = 0; it < 12; it++) {
        chol_small(A, n, lam, Lc);
        tri_solve(Lc, n, b, z);
        zz = 0.f;
        for (int i = 0; i < n; i++) { zz += z[i] * z[i]; mz[i] = -z[i]; }
        if (zz - r * r < 1e-6f * fmaxf(1e-12f, r * r)) break;
        tri_solve(Lc, n, mz, w);
        float zw = 0.f;
        for (int i = 0; i < n; i++) zw += z[i] * w[i];
        float nz = sqrtf(zz);
        lam = fmaxf(0.f, lam + (nz - r) / r * zz / fmaxf(zw, AV_MINVAL));
    }
    if (zz > r * r) { float s = r * rsqrtf(zz); for (int i = 0; i < n; i++) z[i] *= s; }
    for (int i = 0; i < n; i++) y[i] = z[i] * mu[i];
}

// acc += MinvJT_c^T df for contact block c (lane layout as in stage_rows_contact)
__device__ __forceinline__ void apply_block(EnvS &S, const DevModel &m, const float *blk, int tr, int lane, const float *df) {
    int col = lane & 15, half = lane >> 4;
    const float *MJ = blk + 6 * AV_JW;
    float s = MJ[(3 * half) * AV_JW + col] * df[3 * half] + MJ[(3 * half + 1) * AV_JW + col] * df[3 * half + 1] +
              MJ[(3 * half + 2) * AV_JW + col] * df[3 * half + 2];
    s += __shfl_xor_sync(AV_FULL, s, 16);
    int t = col < 8 ? (tr & 0xff) : ((tr >> 8) & 0xff), dl = col & 7;
    if (half == 0 && t != 0xff && dl < m.tree_dofnum[t]) S.acc[m.tree_dofadr[t] + dl] += s;
}
// residual J_c . acc for all 6 rows, broadcast to every lane
__device__ __forceinline__ void block_residual(const EnvS &S, const DevModel &m, const float *blk, int tr, int lane, float *res) {
    int col = lane & 15, half = lane >> 4;
    int t = col < 8 ? (tr & 0xff) : ((tr >> 8) & 0xff), dl = col & 7;
    float x = (t != 0xff && dl < m.tree_dofnum[t]) ? S.acc[m.tree_dofadr[t] + dl] : 0.f;
    float r0, r1, r2;
    block_dot(blk[(3 * half) * AV_JW + col], blk[(3 * half + 1) * AV_JW + col], blk[(3 * half + 2) * AV_JW + col], x, r0, r1, r2);
    res[0] = __shfl_sync(AV_FULL, r0, 0); res[1] = __shfl_sync(AV_FULL, r1, 0); res[2] = __shfl_sync(AV_FULL, r2, 0);
    res[3] = __shfl_sync(AV_FULL, r0, 16); res[4] = __shfl_sync(AV_FULL, r1, 16); res[5] = __shfl_sync(AV_FULL, r2, 16);
}

__device__ inline void stage_solve(const DevModel &m, EnvS &S, float *scratch, int lane, int iters, int noslip_iters) {
    // acc <- M^-1 J^T f_warm  (constraint part of the acceleration); dual cost of the warm start
    for (int i = lane; i < AV_NVP; i += 32) S.acc[i] = 0.f;
    __syncwarp();
    for (int r = 0; r < S.nsc; r++) {
        float f = S.sc_f[r];
        int t = S.sc_tree[r];
        if (lane < m.tree_dofnum[t]) S.acc[m.tree_dofadr[t] + lane] += S.sc_MJ[r * AV_TD + lane] * f;
        __syncwarp();
    }
    for (int c = 0; c < S.ncon; c++) {
        if ((S.c_info[c] >> 20) & 1) continue;
        apply_block(S, m, scratch + c * AV_CBLK, S.c_tree[c], lane, S.c_f + 6 * c);
        __syncwarp();
    }
    float cost = 0.f;
    for (int r = 0; r < S.nsc; r++) {
        float f = S.sc_f[r];
        float ja = S.sc_c1[r] * S.acc[S.sc_dof1[r]] + (S.sc_dof2[r] >= 0 ? S.sc_c2[r] * S.acc[S.sc_dof2[r]] : 0.f);
        cost += f * (0.5f * (ja + S.sc_R[r] * f) + S.sc_b[r]);
    }
    for (int c = 0; c < S.ncon; c++) {
        int info = S.c_info[c], dim = (info >> 16) & 0xf;
        if ((info >> 20) & 1) continue;
        float res[6];
        block_residual(S, m, scratch + c * AV_CBLK, S.c_tree[c], lane, res);
        float Rn = S.c_Rn[c], mu0 = S.c_mu[3 * c], mu1 = S.c_mu[3 * c + 1], mu2 = S.c_mu[3 * c + 2], Rf = Rn / m.impratio;
        float R[6] = {Rn, Rf, Rf, Rf * mu0 * mu0 / (mu1 * mu1), Rf * mu0 * mu0 / (mu2 * mu2), Rf * mu0 * mu0 / (mu2 * mu2)};
        for (int k = 0; k < dim; k++) {
            float f = S.c_f[6 * c + k];
            cost += f * (0.5f * (res[k] + R[k] * f) + S.c_b[6 * c + k]);
        }
    }
    __syncwarp();
    if (cost >= 0.f) {  // the warm start does not beat f = 0
        for (int i = lane; i < AV_NVP; i += 32) S.acc[i] = 0.f;
        for (int i = lane; i < AV_NSC; i += 32) S.sc_f[i] = 0.f;
        for (int i = lane; i < AV_NCON * 6; i += 32) S.c_f[i] = 0.f;
        __syncwarp();
    }
    for (int it = 0; it < iters + noslip_iters; it++) {
        bool noslip = it >= iters;
        // scalar rows: every lane computes the (uniform) update, lanes < nt apply it
        for (int r = 0; r < S.nsc; r++) {
            bool floss = S.sc_lo[r] > -1e37f && S.sc_lo[r] < 0.f;
            if (noslip && !floss) continue;
            float f = S.sc_f[r], R = noslip ? 0.f : S.sc_R[r];
            float ja = S.sc_c1[r] * S.acc[S.sc_dof1[r]] + (S.sc_dof2[r] >= 0 ? S.sc_c2[r] * S.acc[S.sc_dof2[r]] : 0.f);
            float res = S.sc_b[r] + R * f + ja;
            float x = fminf(fmaxf(f - res / (S.sc_A[r] + R), S.sc_lo[r]), S.sc_hi[r]);
            float df = x - f;
            __syncwarp();
            int t = S.sc_tree[r];
            if (lane < m.tree_dofnum[t]) S.acc[m.tree_dofadr[t] + lane] += S.sc_MJ[r * AV_TD + lane] * df;
            if (lane == 0) S.sc_f[r] = x;
            __syncwarp();
        }
        for (int c = 0; c < S.ncon; c++) {
            int info = S.c_info[c], dim = (info >> 16) & 0xf;
            if ((info >> 20) & 1) continue;
            const float *blk = scratch + c * AV_CBLK;
            int tr = S.c_tree[c];
            float res[6], f[6], old[6], AR[21];
            block_residual(S, m, blk, tr, lane, res);
            float Rn = S.c_Rn[c], mu0 = S.c_mu[3 * c], mu1 = S.c_mu[3 * c + 1], mu2 = S.c_mu[3 * c + 2], Rf = Rn / m.impratio;
            float R[6] = {Rn, Rf, Rf, Rf * mu0 * mu0 / (mu1 * mu1), Rf * mu0 * mu0 / (mu2 * mu2), Rf * mu0 * mu0 / (mu2 * mu2)};
            float mu[5] = {mu0, mu0, mu1, mu2, mu2};
            for (int k = 0; k < 21; k++) AR[k] = blk[12 * AV_JW + k];
            for (int k = 0; k < 6; k++) { old[k] = f[k] = S.c_f[6 * c + k]; }
            if (!noslip) {
                for (int k = 0; k < dim; k++) res[k] += S.c_b[6 * c + k] + R[k] * old[k];
                // (a) ray update (normal-only when the block is empty)
                if (old[0] < AV_MINVAL) {
                    f[0] = fmaxf(0.f, old[0] - res[0] / AR[0]);
                    for (int k = 1; k < dim; k++) f[k] = 0.f;
                } else {
                    float vAv = 0.f, vr = 0.f;
                    for (int k = 0; k < dim; k++) {
                        vr += old[k] * res[k];
                        for (int l = 0; l < dim; l++) vAv += old[k] * AR[tri(k, l)] * old[l];
                    }
                    if (vAv > AV_MINVAL) {
                        float x = -vr / vAv;
                        if (old[0] + x * old[0] < 0.f) x = -1.f;
                        for (int k = 0; k < dim; k++) f[k] = old[k] + x * old[k];
                    }
                }
                // (b) friction rows on the ellipsoid of radius f_n
                if (f[0] < AV_MINVAL) {
                    for (int k = 1; k < dim; k++) f[k] = 0.f;
                } else {
                    float Ac[5][5], bc[5], y[5];
                    for (int k = 1; k < dim; k++) {
                        float s = res[k] + AR[tri(k, 0)] * (f[0] - old[0]);
                        for (int l = 1; l < dim; l++) { Ac[k - 1][l - 1] = AR[tri(k, l)]; s -= AR[tri(k, l)] * old[l]; }
                        bc[k - 1] = s;
                    }
                    qcqp(dim - 1, Ac, bc, mu, f[0], blk + 12 * AV_JW + 21, y);
                    for (int k = 1; k < dim; k++) f[k] = y[k - 1];
                }
            } else {
                // noslip: friction rows only, unregularised A, normal force fixed
                if (old[0] < AV_MINVAL) {
                    for (int k = 1; k < dim; k++) f[k] = 0.f;
                } else {
                    float Ac[5][5], bc[5], y[5];
                    for (int k = 1; k < dim; k++) {
                        float s = res[k] + S.c_b[6 * c + k];
                        for (int l = 1; l < dim; l++) {
                            float a = AR[tri(k, l)] - (k == l ? R[k] : 0.f);
                            Ac[k - 1][l - 1] = a;
                            s -= a * old[l];
                        }
                        bc[k - 1] = s;
                    }
                    qcqp(dim - 1, Ac, bc, mu, old[0], nullptr, y);
                    for (int k = 1; k < dim; k++) f[k] = y[k - 1];
                }
            }
            float df[6];
            for (int k = 0; k < 6; k++) df[k] = (k < dim) ? f[k] - old[k] : 0.f;
            __syncwarp();
            apply_block(S, m, blk, tr, lane, df);
            if (lane < 6) S.c_f[6 * c + lane] = (lane < dim) ? f[lane] : 0.f;
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ K7: Euler with implicit joint damping
__device__ inline void stage_integrate(const DevModel &m, EnvS &S, int lane) {
    float h = m.timestep;
    // total generalized force = smooth + constraint = smooth + M * acc   (acc = M^-1 J^T f)
    float tot = 0.f;
    int i = lane, i2 = lane + 32;
    float tot2 = 0.f;
    for (int pass = 0; pass < 2; pass++) {
        int d = pass ? i2 : i;
        if (d < m.nv) {
            int t = m.dof_tree[d], d0 = m.tree_dofadr[t], nt = m.tree_dofnum[t];
            const float *row = S.M + t * AV_TD * AV_TD + (d - d0) * AV_TD;
            float s = S.qfrc_smooth[d];
            for (int k = 0; k < nt; k++) s += row[k] * S.acc[d0 + k];
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
        chol_block(S.M + lane * AV_TD * AV_TD, Lb, nt, m.dof_damping + d0, h);
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
__device__ inline int stage_reward(const DevModel &m, const EnvS &S, int lane, int &latch) {
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
