// avsim_model_pack.h -- host-side: read a compiled .avm model, derive the run-time tables and pack everything into
// one fp32 blob + one int32 blob whose DevModel pointers are relocated to wherever the blobs end up (device memory
// for libavsim.so, host memory for the warp-emulation debug harness under tests/emu).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "avsim_dev.h"

// ------------------------------------------------------------------ .avm reader
struct AvmEntry {
    char name[32];
    uint32_t dtype, ndim, shape[4];
    uint64_t off, nbytes;
};
struct Avm {
    std::vector<uint8_t> blob;
    std::map<std::string, const AvmEntry *> toc;
    bool load(const char *path) {
        FILE *fh = fopen(path, "rb");
        if (!fh) return false;
        fseek(fh, 0, SEEK_END);
        long sz = ftell(fh);
        fseek(fh, 0, SEEK_SET);
        blob.resize(sz);
        bool ok = fread(blob.data(), 1, sz, fh) == (size_t)sz && sz > 12 && !memcmp(blob.data(), "AVSIMMD1", 8);
        fclose(fh);
        if (!ok) return false;
        uint32_t n = *(uint32_t *)(blob.data() + 8);
        const AvmEntry *e = (const AvmEntry *)(blob.data() + 12);
        for (uint32_t i = 0; i < n; i++) toc[std::string(e[i].name, strnlen(e[i].name, 32))] = &e[i];
        return true;
    }
    bool has(const char *n) const { return toc.count(n) != 0; }
    size_t count(const char *n) const {
        const AvmEntry *e = toc.at(n);
        size_t c = 1;
        for (uint32_t k = 0; k < e->ndim; k++) c *= e->shape[k];
        return c;
    }
    int len(const char *n) const { return (int)toc.at(n)->shape[0]; }
    std::vector<float> f(const char *n) const {
        const AvmEntry *e = toc.at(n);
        size_t c = count(n);
        std::vector<float> out(c);
        if (e->dtype == 0) { const double *p = (const double *)(blob.data() + e->off); for (size_t i = 0; i < c; i++) out[i] = (float)p[i]; }
        else { const int32_t *p = (const int32_t *)(blob.data() + e->off); for (size_t i = 0; i < c; i++) out[i] = (float)p[i]; }
        return out;
    }
    std::vector<double> d(const char *n) const {
        const AvmEntry *e = toc.at(n);
        size_t c = count(n);
        std::vector<double> out(c);
        const double *p = (const double *)(blob.data() + e->off);
        for (size_t i = 0; i < c; i++) out[i] = p[i];
        return out;
    }
    std::vector<int> i(const char *n) const {
        const AvmEntry *e = toc.at(n);
        size_t c = count(n);
        std::vector<int> out(c);
        const int32_t *p = (const int32_t *)(blob.data() + e->off);
        for (size_t k = 0; k < c; k++) out[k] = p[k];
        return out;
    }
};

namespace avpack {
struct Packer {
    std::vector<float> fdata;
    std::vector<int> idata;
    std::vector<std::pair<const float **, size_t>> fptr;
    std::vector<std::pair<const int **, size_t>> iptr;
    void addf(const float **dst, const std::vector<float> &v) {
        while (fdata.size() % 4) fdata.push_back(0.f);
        fptr.push_back({dst, fdata.size()});
        fdata.insert(fdata.end(), v.begin(), v.end());
        if (v.empty()) fdata.push_back(0.f);
    }
    void addi(const int **dst, const std::vector<int> &v) {
        while (idata.size() % 4) idata.push_back(0);
        iptr.push_back({dst, idata.size()});
        idata.insert(idata.end(), v.begin(), v.end());
        if (v.empty()) idata.push_back(0);
    }
};
void quat2mat_d(const double *q, double *m) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
void quat_mul_d(const double *a, const double *b, double *r) {
    r[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    r[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    r[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    r[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}


struct PackedModel {
    DevModel dm;
    Packer P;
    std::vector<float> hull4;   // xyzw per hull vertex
    std::string error;
    // point the DevModel at the final location of the blobs
    void relocate(const float *fbase, const int *ibase, const void *hullbase) {
        for (auto &pr : P.fptr) *pr.first = fbase + pr.second;
        for (auto &pr : P.iptr) *pr.first = ibase + pr.second;
        dm.hull_vert = (const float4 *)hullbase;
    }
};

inline bool pack_model(const char *avm_path, PackedModel &out) {
    Avm a;
    if (!avm_path || !a.load(avm_path)) { out.error = "cannot read model file"; return false; }
    DevModel &d = out.dm;
    memset(&d, 0, sizeof d);
    d.nbody = a.len("body_parent"); d.njnt = a.len("jnt_type"); d.nv = a.len("dof_body"); d.nq = a.len("qpos0");
    d.ngeom = a.len("geom_type"); d.npair = a.len("pair_geom"); d.nu = a.len("act_dof"); d.neq = a.len("eq_dof1");
    d.nfree = a.len("free_qadr");
    d.task_id = a.i("task_id")[0]; d.max_reward = a.i("max_reward")[0]; d.num_arms = a.i("num_arms")[0];
    d.noslip_iterations = a.i("noslip_iterations")[0]; d.multiccd = a.i("multiccd")[0];
    d.nj_obs = d.num_arms == 3 ? 21 : 14;
    d.timestep = (float)a.d("timestep")[0]; d.impratio = (float)a.d("impratio")[0];
    auto grav = a.d("gravity");
    for (int k = 0; k < 3; k++) d.gravity[k] = (float)grav[k];
    if (d.nbody > AV_NB || d.nv > AV_NV || d.nq > AV_NQ || d.ngeom > AV_NG || d.nu > AV_NU) {
        out.error = "model exceeds the compiled-in size limits";
        return false;
    }
    Packer &P = out.P;
    auto body_parent = a.i("body_parent"), body_tree = a.i("body_tree"), body_dofadr = a.i("body_dofadr"),
         body_dofnum = a.i("body_dofnum"), dof_body = a.i("dof_body");
    // derived: last dof on the path to each body, tree tables, dof -> tree
    std::vector<int> lastdof(d.nbody, -1);
    for (int b = 1; b < d.nbody; b++) lastdof[b] = body_dofnum[b] ? body_dofadr[b] + body_dofnum[b] - 1 : lastdof[body_parent[b]];
    int ntree = 0;
    for (int b = 0; b < d.nbody; b++) ntree = std::max(ntree, body_tree[b] + 1);
    d.ntree = ntree;
    std::vector<int> tb0(ntree, 1 << 30), tbn(ntree, 0), td0(ntree, 1 << 30), tdn(ntree, 0), dof_tree(d.nv);
    for (int b = 0; b < d.nbody; b++)
        if (body_tree[b] >= 0) { tb0[body_tree[b]] = std::min(tb0[body_tree[b]], b); tbn[body_tree[b]]++; }
    for (int i = 0; i < d.nv; i++) {
        int t = body_tree[dof_body[i]];
        dof_tree[i] = t; td0[t] = std::min(td0[t], i); tdn[t]++;
    }
    for (int t = 0; t < ntree; t++)
        if (tdn[t] > AV_TD || ntree > AV_NTREE) { out.error = "kinematic tree too large"; return false; }
    auto dof_parent_v = a.i("dof_parent");
    std::vector<int> treemask(d.nbody, 0);
    for (int b = 1; b < d.nbody; b++)
        if (body_tree[b] >= 0)
            for (int i = lastdof[b]; i >= 0; i = dof_parent_v[i]) treemask[b] |= 1 << (i - td0[body_tree[b]]);
    // static world poses (exact for world-welded bodies; compile-time pose otherwise)
    auto bpos = a.d("body_pos"), bquat = a.d("body_quat");
    std::vector<double> xpos(3 * d.nbody, 0.0), xquat(4 * d.nbody, 0.0);
    xquat[0] = 1;
    for (int b = 1; b < d.nbody; b++) {
        int p = body_parent[b];
        double R[9];
        quat2mat_d(&xquat[4 * p], R);
        for (int r = 0; r < 3; r++) xpos[3 * b + r] = xpos[3 * p + r] + R[3 * r] * bpos[3 * b] + R[3 * r + 1] * bpos[3 * b + 1] + R[3 * r + 2] * bpos[3 * b + 2];
        quat_mul_d(&xquat[4 * p], &bquat[4 * b], &xquat[4 * b]);
    }
    auto geom_body = a.i("geom_body"), geom_type = a.i("geom_type");
    auto gpos = a.d("geom_pos"), gquat = a.d("geom_quat");
    std::vector<float> gmat(9 * d.ngeom), gx0(3 * d.ngeom), gm0(9 * d.ngeom), ga0(3 * d.ngeom);
    auto gaabb = a.d("geom_aabb");
    std::vector<int> gstatic(d.ngeom);
    for (int g = 0; g < d.ngeom; g++) {
        int b = geom_body[g];
        double Rl[9], Rb[9];
        quat2mat_d(&gquat[4 * g], Rl);
        quat2mat_d(&xquat[4 * b], Rb);
        gstatic[g] = body_tree[b] < 0;
        for (int r = 0; r < 3; r++) {
            double h = 0;
            for (int c = 0; c < 3; c++) {
                double e = Rb[3 * r] * Rl[c] + Rb[3 * r + 1] * Rl[3 + c] + Rb[3 * r + 2] * Rl[6 + c];
                h += (e < 0 ? -e : e) * gaabb[3 * g + c];
            }
            ga0[3 * g + r] = (float)h;
        }
        for (int r = 0; r < 3; r++) {
            gx0[3 * g + r] = (float)(xpos[3 * b + r] + Rb[3 * r] * gpos[3 * g] + Rb[3 * r + 1] * gpos[3 * g + 1] + Rb[3 * r + 2] * gpos[3 * g + 2]);
            for (int c = 0; c < 3; c++) {
                gmat[9 * g + 3 * r + c] = (float)Rl[3 * r + c];
                gm0[9 * g + 3 * r + c] = (float)(Rb[3 * r] * Rl[c] + Rb[3 * r + 1] * Rl[3 + c] + Rb[3 * r + 2] * Rl[6 + c]);
            }
        }
    }
    // pair table: g1 | g2 << 8 | narrowphase class << 16, bounding-sphere radius sums
    auto pair = a.i("pair_geom");
    auto rb = a.f("geom_rbound");
    std::vector<int> pk(d.npair);
    std::vector<float> rsum(d.npair);
    for (int p = 0; p < d.npair; p++) {
        int g1 = pair[2 * p], g2 = pair[2 * p + 1], t1 = geom_type[g1], t2 = geom_type[g2], ty = AV_PAIR_CONVEX;
        if (t1 == AV_GEOM_SPHERE && t2 == AV_GEOM_SPHERE) ty = AV_PAIR_SS;
        else if (t1 == AV_GEOM_SPHERE && t2 == AV_GEOM_BOX) ty = AV_PAIR_SB;
        else if (t1 == AV_GEOM_BOX && t2 == AV_GEOM_SPHERE) ty = AV_PAIR_BS;
        else if (t1 == AV_GEOM_BOX && t2 == AV_GEOM_BOX) ty = AV_PAIR_BB;
        pk[p] = g1 | (g2 << 8) | (ty << 16);
        rsum[p] = rb[g1] + rb[g2];
    }
    std::vector<float> xposf(xpos.begin(), xpos.end()), xquatf(xquat.begin(), xquat.end());

#define PF(field, name) P.addf(&d.field, a.f(name))
#define PI(field, name) P.addi(&d.field, a.i(name))
    PI(body_parent, "body_parent"); PI(body_jntadr, "body_jntadr"); PI(body_jntnum, "body_jntnum");
    PI(body_dofadr, "body_dofadr"); PI(body_dofnum, "body_dofnum"); PI(body_tree, "body_tree");
    P.addi(&d.body_lastdof, lastdof); P.addi(&d.body_treemask, treemask);
    PF(body_pos, "body_pos"); PF(body_quat, "body_quat"); PF(body_mass, "body_mass"); PF(body_ipos, "body_ipos");
    PF(body_inertia, "body_inertia"); PF(body_invweight0, "body_invweight0");
    P.addf(&d.body_xpos0, xposf); P.addf(&d.body_xquat0, xquatf);
    P.addi(&d.tree_bodyadr, tb0); P.addi(&d.tree_bodynum, tbn); P.addi(&d.tree_dofadr, td0); P.addi(&d.tree_dofnum, tdn);
    PI(jnt_type, "jnt_type"); PI(jnt_qposadr, "jnt_qposadr"); PI(jnt_dofadr, "jnt_dofadr"); PI(jnt_limited, "jnt_limited");
    PF(jnt_axis, "jnt_axis"); PF(jnt_pos, "jnt_pos"); PF(jnt_range, "jnt_range"); PF(jnt_solref, "jnt_solref");
    PF(jnt_solimp, "jnt_solimp");
    PI(dof_body, "dof_body"); PI(dof_jnt, "dof_jnt"); PI(dof_parent, "dof_parent"); P.addi(&d.dof_tree, dof_tree);
    PI(dof_frc_limited, "dof_frc_limited");
    PF(dof_armature, "dof_armature"); PF(dof_damping, "dof_damping"); PF(dof_frictionloss, "dof_frictionloss");
    PF(dof_frc_lo, "dof_frc_lo"); PF(dof_frc_hi, "dof_frc_hi"); PF(dof_invweight0, "dof_invweight0");
    PF(dof_solref, "dof_solref"); PF(dof_solimp, "dof_solimp"); PF(qpos0, "qpos0");
    PI(geom_type, "geom_type"); PI(geom_body, "geom_body"); PI(geom_condim, "geom_condim"); PI(geom_hull, "geom_hull");
    PI(geom_class, "geom_class"); P.addi(&d.geom_static, gstatic);
    PF(geom_pos, "geom_pos"); P.addf(&d.geom_mat, gmat); PF(geom_size, "geom_size"); PF(geom_rbound, "geom_rbound");
    PF(geom_aabb, "geom_aabb"); PF(geom_friction, "geom_friction"); PF(geom_solref, "geom_solref");
    PF(geom_solimp, "geom_solimp"); PF(geom_gap, "geom_gap"); PF(geom_margin, "geom_margin");
    P.addf(&d.geom_xpos0, gx0); P.addf(&d.geom_xmat0, gm0); P.addf(&d.geom_xaabb0, ga0);
    PI(hull_adr, "hull_adr"); PI(hull_num, "hull_num");
    P.addi(&d.pair_geom, pk); P.addf(&d.pair_rsum, rsum);
    PI(eq_dof1, "eq_dof1"); PI(eq_dof2, "eq_dof2"); PI(eq_qadr1, "eq_qadr1"); PI(eq_qadr2, "eq_qadr2");
    PF(eq_polycoef, "eq_polycoef"); PF(eq_solref, "eq_solref"); PF(eq_solimp, "eq_solimp"); PF(eq_invweight0, "eq_invweight0");
    PI(act_dof, "act_dof"); PI(act_qadr, "act_qadr");
    PF(act_kp, "act_kp"); PF(act_kv, "act_kv"); PF(act_ctrl_lo, "act_ctrl_lo"); PF(act_ctrl_hi, "act_ctrl_hi");
    PI(obs_qadr, "obs_qadr"); PI(finger_qadr, "finger_qadr"); PI(free_qadr, "free_qadr");
    PF(reset_lo, "reset_lo"); PF(reset_hi, "reset_hi"); PI(reset_draw, "reset_draw");
    d.ncam = a.len("cam_body");
    PI(cam_body, "cam_body"); PI(geom_visible, "geom_visible");
    PF(cam_pos, "cam_pos"); PF(cam_quat, "cam_quat"); PF(cam_fovy, "cam_fovy"); PF(geom_rgba, "geom_rgba");
    PI(geom_tex, "geom_tex"); PF(hull_kdop, "hull_kdop"); PF(light, "light"); PF(table_tex, "table_tex");
    PI(ik_ndof, "ik_ndof"); PF(ik_w0, "ik_w0"); PF(ik_p0, "ik_p0"); PF(ik_site0, "ik_site0"); PF(ik_range, "ik_range");
#undef PF
#undef PI
    auto hv = a.f("hull_vert");
    size_t nh = hv.size() / 3;
    out.hull4.assign(4 * (nh ? nh : 1), 0.f);
    for (size_t i = 0; i < nh; i++) { out.hull4[4 * i] = hv[3 * i]; out.hull4[4 * i + 1] = hv[3 * i + 1]; out.hull4[4 * i + 2] = hv[3 * i + 2]; }
    return true;
}
}  // namespace avpack

// home pose (reference constants.py:26-28)
static const float AV_HOME[21] = {0, -0.082f, 1.06f, 0, -0.953f, 0, 0.02239f, 0, -0.082f, 1.06f, 0, -0.953f, 0, 0.02239f,
                                  0, -0.8f, 0.8f, 0, 0.5f, 0, 0};
