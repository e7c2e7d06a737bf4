/*
 * avsim_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU, fp64, single environment, scalar).
 *
 * A plain-C restatement of the physics step the reference executes through
 *   GuidedVisionEnv.step  (reference gym_guided_vision/gym_guided_vision/env.py:203-226)
 *   -> Physics.step(nstep=20) (env.py:218) -> MuJoCo mj_step  [third-party, NOT in the reference tree]
 * with the options of reference assets/aloha_sim.xml:4-6 (elliptic cones, impratio 100, noslip 3).
 *
 * PARITY UNPINNED: MuJoCo (mujoco ^3.2.2, gym_guided_vision/pyproject.toml:10) cannot be installed in
 * the build container and the reference tree holds no golden qpos/qvel/contact vectors, so this file
 * restates MuJoCo's *published* pipeline and is validated by closed forms and
 * invariants (tests/test_physics_known_answers.py: free fall + spin, static equilibrium, box inertias, virtual work,
 * kinetic energy, cone feasibility, box-box / sphere-box / MPR narrowphase), not against MuJoCo outputs.  Where the algorithm is a free choice
 * (box-box manifold, MPR penetration after libccd), the choice is stated here and the CUDA path (av_aloha_b200/csrc) follows the
 * same statement in fp32, written independently.  The constraint solve is MuJoCo's default: Newton on the primal problem with an
 * exact line search (solve_newton; tolerance-based, + noslip sweeps); the block Gauss-Seidel on the dual of round 1 is kept as a
 * second, independent algorithm for the same optimum (ora_set_solver) -- the two are checked against each other
 * (tests/test_solver_newton.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NB_MAX 40
#define NJ_MAX 40
#define NV_MAX 48
#define NQ_MAX 56
#define NG_MAX 100
#define NU_MAX 24
#define NCON_MAX 256
#define NEFC_MAX (32 + 6 * NCON_MAX)
#define MINVAL 1e-15

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_SPHERE = 2, GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { ROW_EQ = 0, ROW_FLOSS = 1, ROW_LIMIT = 2, ROW_CONTACT = 3 };

/* ------------------------------------------------------------------ model (.avm reader) */
typedef struct {
    char name[32];
    uint32_t dtype, ndim, shape[4];
    uint64_t off, nbytes;
} avm_entry;

typedef struct ora_model {
    uint8_t *blob;
    uint32_t narr;
    avm_entry *toc;
    int nbody, njnt, nv, nq, ngeom, npair, nu, neq, nhull, nfree;
    int task_id, max_reward, num_arms, noslip_iterations, multiccd;
    double timestep, impratio;
    const double *gravity;
    const int *body_parent, *body_jntadr, *body_jntnum, *body_dofadr, *body_dofnum, *body_weld, *body_tree;
    const double *body_pos, *body_quat, *body_mass, *body_ipos, *body_inertia, *body_invweight0;
    const int *jnt_type, *jnt_body, *jnt_qposadr, *jnt_dofadr, *jnt_limited;
    const double *jnt_axis, *jnt_pos, *jnt_range, *jnt_solref, *jnt_solimp;
    const int *dof_body, *dof_jnt, *dof_parent, *dof_frc_limited;
    const double *dof_armature, *dof_damping, *dof_frictionloss, *dof_frc_lo, *dof_frc_hi, *dof_invweight0,
        *dof_solref, *dof_solimp;
    const double *qpos0;
    const int *geom_type, *geom_body, *geom_condim, *geom_hull, *geom_class;
    const double *geom_pos, *geom_quat, *geom_size, *geom_rbound, *geom_aabb, *geom_friction,
        *geom_solref, *geom_solimp, *geom_gap, *geom_margin;
    const int *hull_adr, *hull_num;
    const double *hull_vert;
    const int *pair_geom;
    const int *eq_dof1, *eq_dof2, *eq_qadr1, *eq_qadr2;
    const double *eq_polycoef, *eq_solref, *eq_solimp, *eq_invweight0;
    const int *act_dof, *act_qadr;
    const double *act_kp, *act_kv, *act_ctrl_lo, *act_ctrl_hi;
    const int *obs_qadr, *finger_qadr, *free_qadr;
    int body_lastdof[NB_MAX];
} ora_model;

static const avm_entry *avm_find(const ora_model *m, const char *name) {
    for (uint32_t i = 0; i < m->narr; i++)
        if (!strncmp(m->toc[i].name, name, 32)) return &m->toc[i];
    fprintf(stderr, "avsim_oracle: array '%s' missing from model\n", name);
    abort();
}
static const double *avm_f(const ora_model *m, const char *n) { return (const double *)(m->blob + avm_find(m, n)->off); }
static const int *avm_i(const ora_model *m, const char *n) { return (const int *)(m->blob + avm_find(m, n)->off); }
static int avm_len(const ora_model *m, const char *n) { return (int)avm_find(m, n)->shape[0]; }

ora_model *ora_model_load(const char *path) {
    FILE *fh = fopen(path, "rb");
    if (!fh) return NULL;
    fseek(fh, 0, SEEK_END);
    long sz = ftell(fh);
    fseek(fh, 0, SEEK_SET);
    ora_model *m = (ora_model *)calloc(1, sizeof(ora_model));
    m->blob = (uint8_t *)malloc(sz);
    if (fread(m->blob, 1, sz, fh) != (size_t)sz || memcmp(m->blob, "AVSIMMD1", 8)) {
        fclose(fh);
        free(m->blob);
        free(m);
        return NULL;
    }
    fclose(fh);
    m->narr = *(uint32_t *)(m->blob + 8);
    m->toc = (avm_entry *)(m->blob + 12);
#define F(x) m->x = avm_f(m, #x)
#define I(x) m->x = avm_i(m, #x)
    m->nbody = avm_len(m, "body_parent"); m->njnt = avm_len(m, "jnt_type"); m->nv = avm_len(m, "dof_body");
    m->nq = avm_len(m, "qpos0"); m->ngeom = avm_len(m, "geom_type"); m->npair = avm_len(m, "pair_geom");
    m->nu = avm_len(m, "act_dof"); m->neq = avm_len(m, "eq_dof1"); m->nhull = avm_len(m, "hull_adr");
    m->nfree = avm_len(m, "free_qadr");
    m->task_id = avm_i(m, "task_id")[0]; m->max_reward = avm_i(m, "max_reward")[0];
    m->num_arms = avm_i(m, "num_arms")[0]; m->noslip_iterations = avm_i(m, "noslip_iterations")[0];
    m->multiccd = avm_i(m, "multiccd")[0];
    m->timestep = avm_f(m, "timestep")[0]; m->impratio = avm_f(m, "impratio")[0];
    F(gravity); I(body_parent); I(body_jntadr); I(body_jntnum); I(body_dofadr); I(body_dofnum); I(body_weld); I(body_tree);
    F(body_pos); F(body_quat); F(body_mass); F(body_ipos); F(body_inertia); F(body_invweight0);
    I(jnt_type); I(jnt_body); I(jnt_qposadr); I(jnt_dofadr); I(jnt_limited);
    F(jnt_axis); F(jnt_pos); F(jnt_range); F(jnt_solref); F(jnt_solimp);
    I(dof_body); I(dof_jnt); I(dof_parent); I(dof_frc_limited);
    F(dof_armature); F(dof_damping); F(dof_frictionloss); F(dof_frc_lo); F(dof_frc_hi); F(dof_invweight0);
    F(dof_solref); F(dof_solimp); F(qpos0);
    I(geom_type); I(geom_body); I(geom_condim); I(geom_hull); I(geom_class);
    F(geom_pos); F(geom_quat); F(geom_size); F(geom_rbound); F(geom_aabb); F(geom_friction);
    F(geom_solref); F(geom_solimp); F(geom_gap); F(geom_margin);
    I(hull_adr); I(hull_num); F(hull_vert); I(pair_geom);
    I(eq_dof1); I(eq_dof2); I(eq_qadr1); I(eq_qadr2); F(eq_polycoef); F(eq_solref); F(eq_solimp); F(eq_invweight0);
    I(act_dof); I(act_qadr); F(act_kp); F(act_kv); F(act_ctrl_lo); F(act_ctrl_hi);
    I(obs_qadr); I(finger_qadr); I(free_qadr);
#undef F
#undef I
    if (m->nbody > NB_MAX || m->nv > NV_MAX || m->nq > NQ_MAX || m->ngeom > NG_MAX || m->njnt > NJ_MAX) abort();
    m->body_lastdof[0] = -1;
    for (int b = 1; b < m->nbody; b++)
        m->body_lastdof[b] = m->body_dofnum[b] ? m->body_dofadr[b] + m->body_dofnum[b] - 1
                                               : m->body_lastdof[m->body_parent[b]];
    return m;
}
void ora_model_free(ora_model *m) {
    if (!m) return;
    free(m->blob);
    free(m);
}

/* ------------------------------------------------------------------ small math */
typedef double v3[3];
static inline void v3set(double *r, double x, double y, double z) { r[0] = x; r[1] = y; r[2] = z; }
static inline void v3cpy(double *r, const double *a) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; }
static inline void v3add(double *r, const double *a, const double *b) { r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; }
static inline void v3sub(double *r, const double *a, const double *b) { r[0] = a[0] - b[0]; r[1] = a[1] - b[1]; r[2] = a[2] - b[2]; }
static inline void v3scl(double *r, const double *a, double s) { r[0] = a[0] * s; r[1] = a[1] * s; r[2] = a[2] * s; }
static inline void v3axpy(double *r, double s, const double *a) { r[0] += a[0] * s; r[1] += a[1] * s; r[2] += a[2] * s; }
static inline double v3dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void v3cross(double *r, const double *a, const double *b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline double v3norm(const double *a) { return sqrt(v3dot(a, a)); }
static inline double v3normalize(double *a) {
    double n = v3norm(a);
    if (n > MINVAL) { a[0] /= n; a[1] /= n; a[2] /= n; }
    return n;
}
/* row-major 3x3 */
static inline void m3mulv(double *r, const double *m, const double *v) {
    double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2],
           z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static inline void m3tmulv(double *r, const double *m, const double *v) {
    double x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2], y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2],
           z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
    r[0] = x; r[1] = y; r[2] = z;
}
static void m3mul(double *r, const double *a, const double *b) {
    double t[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
    memcpy(r, t, sizeof t);
}
static void quat_mul(double *r, const double *a, const double *b) {
    double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    double y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    double z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
    r[0] = w; r[1] = x; r[2] = y; r[3] = z;
}
static void quat_normalize(double *q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    for (int i = 0; i < 4; i++) q[i] /= n;
}
static void quat2mat(double *m, const double *q) {
    double w = q[0], x = q[1], y = q[2], z = q[3];
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}

/* ------------------------------------------------------------------ per-environment data */
typedef struct {
    double dist, pos[3], frame[9]; /* frame rows: normal (geom1 -> geom2), tangent1, tangent2 */
    int geom1, geom2, dim, excluded, efc_adr;
    double friction[5], solref[2], solimp[5], includemargin;
} ora_contact;

typedef struct {
    int max_iter;       /* PGS sweeps (oracle default: run to convergence) */
    double tol;         /* stop when the sweep's dual-cost decrease / trace(M) falls below (MuJoCo-style scaling) */
    int noslip_iter;    /* -1: take from model */
    int multiccd;       /* -1: take from model */
    int warmstart;
    int sweep_mode;     /* 0: forward sweeps (what the CUDA kernel does); 1: alternate forward / backward (experiment) */
    int solver;         /* 0: block PGS on the dual; 1: Newton on the primal (the reference's solver, aloha_sim.xml:4-6 leaves MuJoCo's default) */
    int newton_ls;      /* Newton: line-search evaluations per iteration (<= 0: run the line search to convergence) */
} ora_options;

typedef struct ora_data {
    const ora_model *m;
    ora_options opt;
    double qpos[NQ_MAX], qvel[NV_MAX], ctrl[NU_MAX], qacc_warmstart[NV_MAX];
    int latch; /* SewNeedle _threaded_needle (reference env.py:602,631,673) */
    /* position stage */
    double xpos[NB_MAX][3], xquat[NB_MAX][4], xmat[NB_MAX][9], xipos[NB_MAX][3];
    double xanchor[NJ_MAX][3], xaxis[NJ_MAX][3];
    double gpos[NG_MAX][3], gmat[NG_MAX][9];
    double cdof[NV_MAX][6];  /* [ang; lin at world origin] */
    double cinert[NB_MAX][10], crb[NB_MAX][10];
    double M[NV_MAX][NV_MAX], L[NV_MAX][NV_MAX]; /* dense; L = Cholesky factor of M */
    int ncon;
    ora_contact con[NCON_MAX];
    /* velocity / force stage */
    double cvel[NB_MAX][6], cdof_dot[NV_MAX][6];
    double qfrc_bias[NV_MAX], qfrc_passive[NV_MAX], qfrc_actuator[NV_MAX], qfrc_smooth[NV_MAX], qacc_smooth[NV_MAX];
    /* constraints */
    int nefc, ne, nf, nl;
    int efc_type[NEFC_MAX], efc_id[NEFC_MAX];
    double efc_J[NEFC_MAX][NV_MAX], efc_pos[NEFC_MAX], efc_margin[NEFC_MAX], efc_R[NEFC_MAX], efc_aref[NEFC_MAX],
        efc_b[NEFC_MAX], efc_force[NEFC_MAX], efc_floss[NEFC_MAX], efc_vel[NEFC_MAX];
    double *A;   /* nefc x nefc, row stride NEFC_MAX: J M^-1 J^T (no R) */
    double *MinvJT; /* nefc x nv */
    double qacc[NV_MAX], qfrc_constraint[NV_MAX];
    int solver_iters;
    int reward;
    /* force cache (opt.warmstart == 2): forces of the previous solve keyed by constraint identity */
    int cache_n, cache_key[NEFC_MAX], cache_frozen;   /* frozen: physics.forward() reads the cache but does not update it */
    double cache_f[NEFC_MAX][6];
    long *pair_hist;   /* optional [npair] histogram of pairs reaching the narrowphase (diagnostics) */
    int inject_n;      /* >= 0: stage_collision takes this contact list instead of running the narrowphase */
    double inject[NCON_MAX * 9];   /* dist, pos 3, normal 3, geom1, geom2 */
} ora_data;

ora_data *ora_data_new(const ora_model *m) {
    ora_data *d = (ora_data *)calloc(1, sizeof(ora_data));
    d->m = m;
    d->A = (double *)calloc((size_t)NEFC_MAX * NEFC_MAX, sizeof(double));
    d->MinvJT = (double *)calloc((size_t)NEFC_MAX * NV_MAX, sizeof(double));
    d->opt.max_iter = 3000; d->opt.tol = 1e-14; d->opt.noslip_iter = -1; d->opt.multiccd = -1; d->opt.warmstart = 1;
    d->opt.solver = 1;   /* Newton, like the reference (aloha_sim.xml:4-6 leaves MuJoCo's default); ora_set_solver(d, 0, 0) selects the dual PGS */
    d->inject_n = -1;
    memcpy(d->qpos, m->qpos0, m->nq * sizeof(double));
    return d;
}
void ora_data_free(ora_data *d) {
    if (!d) return;
    free(d->A);
    free(d->MinvJT);
    free(d->pair_hist);
    free(d);
}
long *ora_pair_hist(ora_data *d) {
    if (!d->pair_hist) d->pair_hist = (long *)calloc((size_t)d->m->npair, sizeof(long));
    return d->pair_hist;
}
void ora_set_sweep_mode(ora_data *d, int mode) { d->opt.sweep_mode = mode; }
/* n < 0 switches back to the narrowphase */
void ora_inject_contacts(ora_data *d, int n, const double *rec) {
    d->inject_n = n > NCON_MAX ? NCON_MAX : n;
    if (n > 0) memcpy(d->inject, rec, (size_t)d->inject_n * 9 * sizeof(double));
}
void ora_set_solver(ora_data *d, int solver, int newton_ls) { d->opt.solver = solver; d->opt.newton_ls = newton_ls; }
void ora_set_options(ora_data *d, int max_iter, double tol, int noslip_iter, int multiccd, int warmstart) {
    d->opt.max_iter = max_iter; d->opt.tol = tol; d->opt.noslip_iter = noslip_iter; d->opt.multiccd = multiccd;
    d->opt.warmstart = warmstart;
}

/* ------------------------------------------------------------------ stage 1: kinematics (mj_kinematics + mj_comPos) */
static void stage_kinematics(ora_data *d) {
    const ora_model *m = d->m;
    v3set(d->xpos[0], 0, 0, 0);
    d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
    quat2mat(d->xmat[0], d->xquat[0]);
    for (int b = 1; b < m->nbody; b++) {
        int p = m->body_parent[b];
        double pos[3], quat[4], tmp[3];
        m3mulv(tmp, d->xmat[p], m->body_pos + 3 * b);
        v3add(pos, d->xpos[p], tmp);
        quat_mul(quat, d->xquat[p], m->body_quat + 4 * b);
        for (int k = 0; k < m->body_jntnum[b]; k++) {
            int j = m->body_jntadr[b] + k, qa = m->jnt_qposadr[j];
            if (m->jnt_type[j] == JNT_FREE) {
                v3cpy(pos, d->qpos + qa);
                memcpy(quat, d->qpos + qa + 3, 4 * sizeof(double));
                quat_normalize(quat);
                v3cpy(d->xanchor[j], pos);
                v3set(d->xaxis[j], 0, 0, 1);
                continue;
            }
            double R[9];
            quat2mat(R, quat);
            m3mulv(d->xaxis[j], R, m->jnt_axis + 3 * j);
            m3mulv(tmp, R, m->jnt_pos + 3 * j);
            v3add(d->xanchor[j], pos, tmp);
            double dq = d->qpos[qa] - m->qpos0[qa];
            if (m->jnt_type[j] == JNT_SLIDE) {
                v3axpy(pos, dq, d->xaxis[j]);
            } else {
                double s = sin(0.5 * dq), rq[4] = {cos(0.5 * dq), s * m->jnt_axis[3 * j], s * m->jnt_axis[3 * j + 1],
                                                   s * m->jnt_axis[3 * j + 2]};
                quat_mul(quat, quat, rq);
                quat2mat(R, quat);
                m3mulv(tmp, R, m->jnt_pos + 3 * j);
                v3sub(pos, d->xanchor[j], tmp);
            }
        }
        quat_normalize(quat);
        v3cpy(d->xpos[b], pos);
        memcpy(d->xquat[b], quat, sizeof quat);
        quat2mat(d->xmat[b], quat);
        m3mulv(tmp, d->xmat[b], m->body_ipos + 3 * b);
        v3add(d->xipos[b], pos, tmp);
    }
    for (int g = 0; g < m->ngeom; g++) {
        int b = m->geom_body[g];
        double tmp[3], gm[9];
        quat2mat(gm, m->geom_quat + 4 * g);
        m3mulv(tmp, d->xmat[b], m->geom_pos + 3 * g);
        v3add(d->gpos[g], d->xpos[b], tmp);
        m3mul(d->gmat[g], d->xmat[b], gm);
    }
    /* motion axes of every dof, expressed at the world origin */
    for (int i = 0; i < m->nv; i++) {
        int j = m->dof_jnt[i], b = m->dof_body[i];
        double *c = d->cdof[i];
        if (m->jnt_type[j] == JNT_FREE) {
            int k = i - m->jnt_dofadr[j];
            if (k < 3) {
                v3set(c, 0, 0, 0); v3set(c + 3, 0, 0, 0); c[3 + k] = 1;
            } else {
                double ax[3] = {d->xmat[b][k - 3], d->xmat[b][3 + k - 3], d->xmat[b][6 + k - 3]};
                v3cpy(c, ax);
                v3cross(c + 3, d->xpos[b], ax);
            }
        } else if (m->jnt_type[j] == JNT_SLIDE) {
            v3set(c, 0, 0, 0);
            v3cpy(c + 3, d->xaxis[j]);
        } else {
            v3cpy(c, d->xaxis[j]);
            v3cross(c + 3, d->xanchor[j], d->xaxis[j]);
        }
    }
    /* spatial inertia of every body about the world origin: {Ixx Iyy Izz Ixy Ixz Iyz, m*c, m} */
    for (int b = 0; b < m->nbody; b++) {
        double *I = d->cinert[b];
        memset(I, 0, 10 * sizeof(double));
        double mass = m->body_mass[b];
        if (b == 0 || mass == 0) continue;
        const double *i6 = m->body_inertia + 6 * b;
        double Ib[9] = {i6[0], i6[3], i6[4], i6[3], i6[1], i6[5], i6[4], i6[5], i6[2]}, t[9], Rt[9], Iw[9];
        const double *R = d->xmat[b];
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rt[3 * r + c] = R[3 * c + r];
        m3mul(t, R, Ib);
        m3mul(Iw, t, Rt);
        const double *c = d->xipos[b];
        double cc = v3dot(c, c);
        I[0] = Iw[0] + mass * (cc - c[0] * c[0]); I[1] = Iw[4] + mass * (cc - c[1] * c[1]);
        I[2] = Iw[8] + mass * (cc - c[2] * c[2]);
        I[3] = Iw[1] - mass * c[0] * c[1]; I[4] = Iw[2] - mass * c[0] * c[2]; I[5] = Iw[5] - mass * c[1] * c[2];
        I[6] = mass * c[0]; I[7] = mass * c[1]; I[8] = mass * c[2]; I[9] = mass;
    }
}

/* spatial inertia (about origin) times motion vector -> force vector [torque; force] */
static void inert_mul(double *f, const double *I, const double *v) {
    const double *w = v, *l = v + 3, *mc = I + 6;
    double t[3];
    f[0] = I[0] * w[0] + I[3] * w[1] + I[4] * w[2];
    f[1] = I[3] * w[0] + I[1] * w[1] + I[5] * w[2];
    f[2] = I[4] * w[0] + I[5] * w[1] + I[2] * w[2];
    v3cross(t, mc, l);
    v3add(f, f, t);
    v3cross(t, mc, w);
    f[3] = I[9] * l[0] - t[0]; f[4] = I[9] * l[1] - t[1]; f[5] = I[9] * l[2] - t[2];
}
static inline double dot6(const double *a, const double *b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
/* motion x motion */
static void cross_motion(double *r, const double *v, const double *s) {
    double a[3], b[3], c[3];
    v3cross(a, v, s);
    v3cross(b, v, s + 3);
    v3cross(c, v + 3, s);
    v3cpy(r, a);
    v3add(r + 3, b, c);
}
/* motion x* force */
static void cross_force(double *r, const double *v, const double *f) {
    double a[3], b[3], c[3];
    v3cross(a, v, f);
    v3cross(b, v + 3, f + 3);
    v3cross(c, v, f + 3);
    v3add(r, a, b);
    v3cpy(r + 3, c);
}

/* ------------------------------------------------------------------ stage 2: CRB + factor (mj_crb, mj_factorM) */
static int chol_factor(int n, double A[NV_MAX][NV_MAX], double L[NV_MAX][NV_MAX]) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i][j];
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            if (i == j) {
                if (s <= 0) return -1;
                L[i][i] = sqrt(s);
            } else
                L[i][j] = s / L[j][j];
        }
    return 0;
}
static void chol_solve(int n, double L[NV_MAX][NV_MAX], double *x) {
    for (int i = 0; i < n; i++) {
        double s = x[i];
        for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
        x[i] = s / L[i][i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k][i] * x[k];
        x[i] = s / L[i][i];
    }
}

static int stage_inertia(ora_data *d) {
    const ora_model *m = d->m;
    memcpy(d->crb, d->cinert, sizeof(d->crb));
    for (int b = m->nbody - 1; b > 0; b--) {
        int p = m->body_parent[b];
        for (int k = 0; k < 10; k++) d->crb[p][k] += d->crb[b][k];
    }
    memset(d->M, 0, sizeof(d->M));
    for (int i = 0; i < m->nv; i++) {
        double f[6];
        inert_mul(f, d->crb[m->dof_body[i]], d->cdof[i]);
        for (int j = i; j >= 0; j = m->dof_parent[j]) d->M[i][j] = d->M[j][i] = dot6(d->cdof[j], f);
        d->M[i][i] += m->dof_armature[i];
    }
    return chol_factor(m->nv, d->M, d->L);
}

/* ------------------------------------------------------------------ Jacobian of a world point fixed to a body */
static void jac_point(const ora_data *d, int body, const double *p, double jp[3][NV_MAX], double jr[3][NV_MAX]) {
    const ora_model *m = d->m;
    for (int r = 0; r < 3; r++) {
        memset(jp[r], 0, m->nv * sizeof(double));
        memset(jr[r], 0, m->nv * sizeof(double));
    }
    for (int i = m->body_lastdof[body]; i >= 0; i = m->dof_parent[i]) {
        double t[3];
        v3cross(t, d->cdof[i], p);
        for (int r = 0; r < 3; r++) {
            jp[r][i] = d->cdof[i][3 + r] + t[r];
            jr[r][i] = d->cdof[i][r];
        }
    }
}

/* ------------------------------------------------------------------ stage 3: collision (mj_collision) */
#include "avsim_oracle_collide.inc"

/* ------------------------------------------------------------------ stage 4: velocity, bias, actuation (mj_fwdVelocity/Actuation/Acceleration) */
static void stage_smooth(ora_data *d) {
    const ora_model *m = d->m;
    int nv = m->nv;
    /* body velocities + time derivative of the dof axes (mj_comVel) */
    memset(d->cvel, 0, sizeof(d->cvel));
    for (int b = 1; b < m->nbody; b++) {
        double v[6];
        memcpy(v, d->cvel[m->body_parent[b]], sizeof v);
        int a = m->body_dofadr[b], n = m->body_dofnum[b];
        if (n == 6) { /* free joint: translational axes are world-fixed, rotational axes ride on the body */
            for (int k = 0; k < 3; k++) {
                memset(d->cdof_dot[a + k], 0, 6 * sizeof(double));
                for (int r = 0; r < 6; r++) v[r] += d->cdof[a + k][r] * d->qvel[a + k];
            }
            for (int k = 3; k < 6; k++) cross_motion(d->cdof_dot[a + k], v, d->cdof[a + k]);
            for (int k = 3; k < 6; k++)
                for (int r = 0; r < 6; r++) v[r] += d->cdof[a + k][r] * d->qvel[a + k];
        } else {
            for (int k = 0; k < n; k++) {
                cross_motion(d->cdof_dot[a + k], v, d->cdof[a + k]);
                for (int r = 0; r < 6; r++) v[r] += d->cdof[a + k][r] * d->qvel[a + k];
            }
        }
        memcpy(d->cvel[b], v, sizeof v);
    }
    /* RNE with qacc = 0: bias = Coriolis + centrifugal + gravity (mj_rne) */
    static __thread double cacc[NB_MAX][6], cfrc[NB_MAX][6];
    memset(cacc, 0, sizeof cacc);
    memset(cfrc, 0, sizeof cfrc);
    v3scl(cacc[0] + 3, m->gravity, -1.0);
    for (int b = 1; b < m->nbody; b++) {
        memcpy(cacc[b], cacc[m->body_parent[b]], 6 * sizeof(double));
        int a = m->body_dofadr[b];
        for (int k = 0; k < m->body_dofnum[b]; k++)
            for (int r = 0; r < 6; r++) cacc[b][r] += d->cdof_dot[a + k][r] * d->qvel[a + k];
        double Ia[6], Iv[6], vIv[6];
        inert_mul(Ia, d->cinert[b], cacc[b]);
        inert_mul(Iv, d->cinert[b], d->cvel[b]);
        cross_force(vIv, d->cvel[b], Iv);
        for (int r = 0; r < 6; r++) cfrc[b][r] = Ia[r] + vIv[r];
    }
    for (int b = m->nbody - 1; b > 0; b--)
        for (int r = 0; r < 6; r++) cfrc[m->body_parent[b]][r] += cfrc[b][r];
    for (int i = 0; i < nv; i++) d->qfrc_bias[i] = dot6(d->cdof[i], cfrc[m->dof_body[i]]);
    /* passive: joint damping */
    for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -m->dof_damping[i] * d->qvel[i];
    /* position servos: kp*(clamp(ctrl) - q) - kv*qd, then the joint-level actuatorfrcrange clamp */
    memset(d->qfrc_actuator, 0, sizeof(d->qfrc_actuator));
    for (int u = 0; u < m->nu; u++) {
        double c = fmin(fmax(d->ctrl[u], m->act_ctrl_lo[u]), m->act_ctrl_hi[u]);
        d->qfrc_actuator[m->act_dof[u]] += m->act_kp[u] * (c - d->qpos[m->act_qadr[u]]) - m->act_kv[u] * d->qvel[m->act_dof[u]];
    }
    for (int i = 0; i < nv; i++)
        if (m->dof_frc_limited[i]) d->qfrc_actuator[i] = fmin(fmax(d->qfrc_actuator[i], m->dof_frc_lo[i]), m->dof_frc_hi[i]);
    for (int i = 0; i < nv; i++) {
        d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i];
        d->qacc_smooth[i] = d->qfrc_smooth[i];
    }
    chol_solve(nv, d->L, d->qacc_smooth);
}

/* ------------------------------------------------------------------ stage 5: constraint rows (mj_makeConstraint) */
static double impedance(const double *solimp, double pos_minus_margin) {
    double d0 = solimp[0], dw = solimp[1], width = solimp[2], mid = solimp[3], power = solimp[4];
    if (width < MINVAL) return 0.5 * (d0 + dw);
    double x = fabs(pos_minus_margin) / width;
    if (x >= 1) return dw;
    if (x <= 0) return d0;
    double y;
    if (power == 1) y = x;
    else if (x <= mid) y = pow(x / mid, power) * mid; /* a*x^p with a = 1/mid^(p-1) */
    else y = 1 - pow((1 - x) / (1 - mid), power) * (1 - mid);
    return d0 + y * (dw - d0);
}
static void kbi(const ora_model *m, const double *solref, const double *solimp, double pos_minus_margin, double *K,
                double *B, double *imp) {
    double tc = fmax(solref[0], 2 * m->timestep), dr = solref[1], dmax = solimp[1];
    *imp = impedance(solimp, pos_minus_margin);
    *B = 2.0 / (dmax * tc);
    *K = 1.0 / (dmax * dmax * tc * tc * dr * dr);
}
static int add_row(ora_data *d, int type, int id) {
    int r = d->nefc++;
    if (r >= NEFC_MAX) abort();
    memset(d->efc_J[r], 0, d->m->nv * sizeof(double));
    d->efc_type[r] = type; d->efc_id[r] = id; d->efc_floss[r] = 0; d->efc_pos[r] = 0; d->efc_margin[r] = 0;
    return r;
}
static double rowdot(const ora_data *d, int r, const double *v) {
    double s = 0;
    for (int i = 0; i < d->m->nv; i++) s += d->efc_J[r][i] * v[i];
    return s;
}

static void stage_constraint_rows(ora_data *d) {
    const ora_model *m = d->m;
    d->nefc = d->ne = d->nf = d->nl = 0;
    double K, B, imp;
    /* joint equalities: (q1 - q1_0) - poly(q2 - q2_0) = 0 */
    for (int e = 0; e < m->neq; e++) {
        int r = add_row(d, ROW_EQ, e);
        const double *c = m->eq_polycoef + 5 * e;
        double x = d->qpos[m->eq_qadr2[e]] - m->qpos0[m->eq_qadr2[e]];
        double poly = c[0] + x * (c[1] + x * (c[2] + x * (c[3] + x * c[4])));
        double dpoly = c[1] + x * (2 * c[2] + x * (3 * c[3] + x * 4 * c[4]));
        d->efc_pos[r] = d->qpos[m->eq_qadr1[e]] - m->qpos0[m->eq_qadr1[e]] - poly;
        d->efc_J[r][m->eq_dof1[e]] = 1;
        d->efc_J[r][m->eq_dof2[e]] = -dpoly;
        kbi(m, m->eq_solref + 2 * e, m->eq_solimp + 5 * e, d->efc_pos[r], &K, &B, &imp);
        d->efc_vel[r] = rowdot(d, r, d->qvel);
        d->efc_aref[r] = -B * d->efc_vel[r] - K * imp * d->efc_pos[r];
        d->efc_R[r] = fmax(MINVAL, (1 - imp) / imp * m->eq_invweight0[e]);
        d->ne++;
    }
    /* dof friction loss */
    for (int i = 0; i < m->nv; i++)
        if (m->dof_frictionloss[i] > 0) {
            int r = add_row(d, ROW_FLOSS, i);
            d->efc_J[r][i] = 1;
            d->efc_floss[r] = m->dof_frictionloss[i];
            kbi(m, m->dof_solref + 2 * i, m->dof_solimp + 5 * i, 0, &K, &B, &imp);
            d->efc_vel[r] = d->qvel[i];
            d->efc_aref[r] = -B * d->efc_vel[r];
            d->efc_R[r] = fmax(MINVAL, (1 - imp) / imp * m->dof_invweight0[i]);
            d->nf++;
        }
    /* joint limits (active only when violated: margin 0) */
    for (int j = 0; j < m->njnt; j++) {
        if (!m->jnt_limited[j] || m->jnt_type[j] == JNT_FREE) continue;
        double q = d->qpos[m->jnt_qposadr[j]];
        for (int side = 0; side < 2; side++) {
            double dist = side ? m->jnt_range[2 * j + 1] - q : q - m->jnt_range[2 * j];
            if (dist >= 0) continue;
            int r = add_row(d, ROW_LIMIT, j);
            d->efc_J[r][m->jnt_dofadr[j]] = side ? -1 : 1;
            d->efc_pos[r] = dist;
            kbi(m, m->jnt_solref + 2 * j, m->jnt_solimp + 5 * j, dist, &K, &B, &imp);
            d->efc_vel[r] = rowdot(d, r, d->qvel);
            d->efc_aref[r] = -B * d->efc_vel[r] - K * imp * dist;
            d->efc_R[r] = fmax(MINVAL, (1 - imp) / imp * m->dof_invweight0[m->jnt_dofadr[j]]);
            d->nl++;
        }
    }
    /* contacts: elliptic cone rows {normal, tangent1, tangent2, torsion, roll1, roll2}[:dim] */
    static __thread double jp1[3][NV_MAX], jr1[3][NV_MAX], jp2[3][NV_MAX], jr2[3][NV_MAX];
    for (int c = 0; c < d->ncon; c++) {
        ora_contact *con = &d->con[c];
        con->efc_adr = -1;
        if (con->excluded) continue;
        int b1 = m->geom_body[con->geom1], b2 = m->geom_body[con->geom2];
        jac_point(d, b1, con->pos, jp1, jr1);
        jac_point(d, b2, con->pos, jp2, jr2);
        kbi(m, con->solref, con->solimp, con->dist, &K, &B, &imp);
        double Rn = fmax(MINVAL, (1 - imp) / imp * (m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2]));
        for (int k = 0; k < con->dim; k++) {
            int r = add_row(d, ROW_CONTACT, c);
            if (k == 0) con->efc_adr = r;
            const double *ax = con->frame + 3 * (k % 3);
            for (int i = 0; i < m->nv; i++) {
                if (k < 3)
                    d->efc_J[r][i] = ax[0] * (jp2[0][i] - jp1[0][i]) + ax[1] * (jp2[1][i] - jp1[1][i]) + ax[2] * (jp2[2][i] - jp1[2][i]);
                else
                    d->efc_J[r][i] = ax[0] * (jr2[0][i] - jr1[0][i]) + ax[1] * (jr2[1][i] - jr1[1][i]) + ax[2] * (jr2[2][i] - jr1[2][i]);
            }
            d->efc_vel[r] = rowdot(d, r, d->qvel);
            if (k == 0) {
                d->efc_pos[r] = con->dist;
                d->efc_aref[r] = -B * d->efc_vel[r] - K * imp * con->dist;
                d->efc_R[r] = Rn;
            } else {
                d->efc_aref[r] = -B * d->efc_vel[r];
                double fr = con->friction[k - 1];
                d->efc_R[r] = Rn / m->impratio * (con->friction[0] * con->friction[0]) / (fr * fr);
            }
        }
    }
}

/* ------------------------------------------------------------------ stage 6: solve (dual PGS on A+R, then noslip) */
#define A_(i, j) d->A[(size_t)(i) * NEFC_MAX + (j)]

/* minimise 0.5 y'Ay + y'b s.t. sum (y_i/mu_i)^2 <= r^2, n <= 5, A SPD */
static void small_solve(int n, const double A[5][5], const double *b, double *x) { /* Cholesky, x = -A^-1 b */
    double L[5][5];
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i][j];
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            L[i][j] = (i == j) ? sqrt(fmax(s, MINVAL)) : s / L[j][j];
        }
    for (int i = 0; i < n; i++) {
        double s = -b[i];
        for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
        x[i] = s / L[i][i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[k][i] * x[k];
        x[i] = s / L[i][i];
    }
}
static void qcqp(int n, const double Ain[5][5], const double *bin, const double *mu, double r, double *y) {
    double A[5][5], b[5], z[5];
    for (int i = 0; i < n; i++) {
        b[i] = bin[i] * mu[i];
        for (int j = 0; j < n; j++) A[i][j] = Ain[i][j] * mu[i] * mu[j];
    }
    small_solve(n, A, b, z);
    double zz = 0;
    for (int i = 0; i < n; i++) zz += z[i] * z[i];
    if (zz > r * r) {
        /* Newton on the multiplier: (A + lam I) z = -b, |z| = r */
        double lam = 0;
        for (int it = 0; it < 40; it++) {
            double Al[5][5], w[5], mz[5];
            memcpy(Al, A, sizeof Al);
            for (int i = 0; i < n; i++) Al[i][i] += lam;
            small_solve(n, Al, b, z);
            zz = 0;
            for (int i = 0; i < n; i++) { zz += z[i] * z[i]; mz[i] = -z[i]; }
            double phi = zz - r * r;
            if (phi < 1e-14 * fmax(1.0, r * r)) break;
            small_solve(n, Al, mz, w); /* w = (A+lam I)^-1 z ;  d|z|^2/dlam = -2 z'w */
            double zw = 0;
            for (int i = 0; i < n; i++) zw += z[i] * w[i];
            /* Newton on 1/|z| - 1/r is better conditioned */
            double nz = sqrt(zz);
            lam += (nz - r) / r * zz / fmax(zw, MINVAL);
            if (lam < 0) lam = 0;
        }
        zz = 0;
        for (int i = 0; i < n; i++) zz += z[i] * z[i];
        if (zz > r * r) { double s = r / sqrt(zz); for (int i = 0; i < n; i++) z[i] *= s; }
    }
    for (int i = 0; i < n; i++) y[i] = z[i] * mu[i];
}

static double dual_cost(const ora_data *d, const double *f) {
    double c = 0;
    for (int i = 0; i < d->nefc; i++) {
        double s = 0;
        for (int j = 0; j < d->nefc; j++) s += A_(i, j) * f[j];
        c += f[i] * (0.5 * (s + d->efc_R[i] * f[i]) + d->efc_b[i]);
    }
    return c;
}


/* ---- primal Newton (MuJoCo's default solver, mj_solPrimal with flg_Newton [third-party; restated from its published
 * description, SURVEY.md Appendix A]).  Unconstrained convex problem in the acceleration a:
 *     minimise  0.5 (a - a_smooth)' M (a - a_smooth)  +  sum_blocks s(J a - aref)
 * where s is the convex conjugate of the dual block problem that the PGS path solves: equality rows quadratic
 * 0.5 D y^2; friction loss Huber (|force| <= frictionloss); limits / frictionless one-sided quadratic; elliptic
 * contacts three zones (top: 0, bottom: quadratic, middle: 0.5 Dm (N - mu T)^2 with N = mu y_0, T = |friction_j y_j|,
 * mu = friction_0 sqrt(R_1 / R_0), Dm = D_0 / (mu^2 (1 + mu^2))).  Same optimum as the dual (strictly convex), which
 * tests/test_solver_newton.py checks numerically against the PGS path run to convergence. */
/* cost of one block at constraint-space value y; writes force = -ds/dy and (optionally) the block Hessian */
static double block_cost(const ora_data *d, int i, int dim, const double *y, double *force, double H[6][6]) {
    int type = d->efc_type[i];
    if (H) memset(H, 0, 36 * sizeof(double));
    if (type != ROW_CONTACT) {
        double D = 1.0 / d->efc_R[i], yy = y[0];
        if (type == ROW_EQ) { force[0] = -D * yy; if (H) H[0][0] = D; return 0.5 * D * yy * yy; }
        if (type == ROW_FLOSS) {
            double fl = d->efc_floss[i], Rf = d->efc_R[i] * fl;
            if (yy <= -Rf) { force[0] = fl; return -0.5 * Rf * fl - fl * yy; }
            if (yy >= Rf) { force[0] = -fl; return -0.5 * Rf * fl + fl * yy; }
            force[0] = -D * yy; if (H) H[0][0] = D; return 0.5 * D * yy * yy;
        }
        if (yy >= 0) { force[0] = 0; return 0; }            /* limit: one-sided */
        force[0] = -D * yy; if (H) H[0][0] = D; return 0.5 * D * yy * yy;
    }
    const ora_contact *con = &d->con[d->efc_id[i]];
    if (dim == 1) {
        double D = 1.0 / d->efc_R[i];
        if (y[0] >= 0) { force[0] = 0; return 0; }
        force[0] = -D * y[0]; if (H) H[0][0] = D; return 0.5 * D * y[0] * y[0];
    }
    double mu = con->friction[0] * sqrt(d->efc_R[i + 1] / d->efc_R[i]);
    double sc[6], U[6], T2 = 0;
    sc[0] = mu;
    for (int k = 1; k < dim; k++) sc[k] = con->friction[k - 1];
    for (int k = 0; k < dim; k++) U[k] = sc[k] * y[k];
    for (int k = 1; k < dim; k++) T2 += U[k] * U[k];
    double N = U[0], T = sqrt(T2);
    if (N >= mu * T || (T <= 0 && N >= 0)) {                 /* top zone: inside the polar cone, no force */
        for (int k = 0; k < dim; k++) force[k] = 0;
        return 0;
    }
    if (mu * N + T <= 0 || (T <= 0 && N < 0)) {              /* bottom zone: inside the cone, plain quadratic */
        double c = 0;
        for (int k = 0; k < dim; k++) {
            double D = 1.0 / d->efc_R[i + k];
            force[k] = -D * y[k]; c += 0.5 * D * y[k] * y[k];
            if (H) H[k][k] = D;
        }
        return c;
    }
    double Dm = 1.0 / (d->efc_R[i] * mu * mu * (1 + mu * mu)), e = N - mu * T;   /* middle zone: on the cone surface */
    force[0] = -Dm * e * mu;
    for (int k = 1; k < dim; k++) force[k] = Dm * e * mu * U[k] / T * sc[k];
    if (H) {
        double HU[6][6];
        HU[0][0] = Dm;
        for (int k = 1; k < dim; k++) HU[0][k] = HU[k][0] = -Dm * mu * U[k] / T;
        for (int k = 1; k < dim; k++)
            for (int l = 1; l < dim; l++)
                HU[k][l] = Dm * (mu * mu * U[k] * U[l] / T2 - e * mu * ((k == l) - U[k] * U[l] / T2) / T);
        for (int k = 0; k < dim; k++)
            for (int l = 0; l < dim; l++) H[k][l] = sc[k] * HU[k][l] * sc[l];
    }
    return 0.5 * Dm * e * e;
}
static int block_dim(const ora_data *d, int i) { return d->efc_type[i] == ROW_CONTACT ? d->con[d->efc_id[i]].dim : 1; }

/* total primal cost at acc = a - a_smooth (jar = J acc + b); fills force and, when H != NULL, the Hessian M + J' Hc J */
static double primal_eval(const ora_data *d, const double *acc, const double *jar, double *force, double (*H)[NV_MAX]) {
    int nv = d->m->nv, n = d->nefc;
    double cost = 0;
    for (int i = 0; i < nv; i++) {
        double s = 0;
        for (int j = 0; j < nv; j++) s += d->M[i][j] * acc[j];
        cost += 0.5 * acc[i] * s;
        if (H) for (int j = 0; j < nv; j++) H[i][j] = d->M[i][j];
    }
    for (int i = 0; i < n;) {
        int dim = block_dim(d, i);
        double Hc[6][6];
        cost += block_cost(d, i, dim, jar + i, force + i, H ? Hc : NULL);
        if (H) {
            int nz[NV_MAX], nnz = 0;   /* columns any row of the block touches (<= 14: two kinematic trees) */
            for (int a = 0; a < nv; a++) {
                int any = 0;
                for (int k = 0; k < dim; k++) any |= d->efc_J[i + k][a] != 0;
                if (any) nz[nnz++] = a;
            }
            for (int k = 0; k < dim; k++)
                for (int l = 0; l < dim; l++) {
                    if (Hc[k][l] == 0) continue;
                    const double *Jk = d->efc_J[i + k], *Jl = d->efc_J[i + l];
                    for (int ia = 0; ia < nnz; ia++) {
                        double t = Jk[nz[ia]] * Hc[k][l];
                        if (t == 0) continue;
                        for (int ic = 0; ic < nnz; ic++) H[nz[ia]][nz[ic]] += t * Jl[nz[ic]];
                    }
                }
        }
        i += dim;
    }
    return cost;
}

static long g_stat_refactor = 0, g_stat_lseval = 0;
long ora_stat_lseval(void) { return g_stat_lseval; }
long ora_stat_refactor(void) { return g_stat_refactor; }
static void solve_newton(ora_data *d, double *f) {
    const ora_model *m = d->m;
    int nv = m->nv, n = d->nefc;
    static __thread double H[NV_MAX][NV_MAX], Lh[NV_MAX][NV_MAX];
    double acc[NV_MAX], jar[NEFC_MAX], grad[NV_MAX], p[NV_MAX], jv[NEFC_MAX], ftmp[NEFC_MAX], Mp[NV_MAX];
    double trM = 0;
    for (int i = 0; i < nv; i++) trM += d->M[i][i];
    double scale = 1.0 / trM;    /* MuJoCo: 1 / (meaninertia * max(1, nv)) */
    /* start: the better of a_smooth (acc = 0) and the warm start (MuJoCo mj_fwdConstraint) */
    memset(acc, 0, sizeof acc);
    for (int r = 0; r < n; r++) jar[r] = d->efc_b[r];
    double cost = primal_eval(d, acc, jar, f, NULL);
    if (d->opt.warmstart) {
        double aw[NV_MAX], jw[NEFC_MAX];
        for (int i = 0; i < nv; i++) aw[i] = d->qacc_warmstart[i] - d->qacc_smooth[i];
        for (int r = 0; r < n; r++) jw[r] = rowdot(d, r, aw) + d->efc_b[r];
        double cw = primal_eval(d, aw, jw, ftmp, NULL);
        if (cw < cost) { cost = cw; memcpy(acc, aw, sizeof aw); memcpy(jar, jw, n * sizeof(double)); }
    }
    for (int it = 0; it < d->opt.max_iter && it < 200; it++) {
        cost = primal_eval(d, acc, jar, f, H);
        for (int i = 0; i < nv; i++) {
            double s = 0;
            for (int j = 0; j < nv; j++) s += d->M[i][j] * acc[j];
            for (int r = 0; r < n; r++) s -= d->efc_J[r][i] * f[r];
            grad[i] = s;
        }
        double gn = 0;
        for (int i = 0; i < nv; i++) gn += grad[i] * grad[i];
        d->solver_iters = it;
        if (scale * sqrt(gn) < (getenv("ORA_TOLG") ? atof(getenv("ORA_TOLG")) : d->opt.tol)) break;
        {   /* experiment (ORA_REFRESH=m): re-factor the Hessian only every m-th iteration, preconditioned CG in between */
            static __thread double gprev[NV_MAX], zprev[NV_MAX], pprev[NV_MAX];
            int refresh = getenv("ORA_REFRESH") ? atoi(getenv("ORA_REFRESH")) : 1;
            if (it % refresh == 0) { if (chol_factor(nv, H, Lh)) break; g_stat_refactor++; }
            double z[NV_MAX];
            for (int i = 0; i < nv; i++) z[i] = grad[i];
            chol_solve(nv, Lh, z);
            double beta = 0;
            if (it % refresh != 0 && getenv("ORA_PCG")) {
                double num = 0, den = 0;
                for (int i = 0; i < nv; i++) { num += z[i] * (grad[i] - gprev[i]); den += zprev[i] * gprev[i]; }
                beta = den > 0 ? fmax(0, num / den) : 0;
            }
            for (int i = 0; i < nv; i++) { p[i] = -z[i] + beta * pprev[i]; gprev[i] = grad[i]; zprev[i] = z[i]; pprev[i] = p[i]; }
        }
        /* stop on the Newton decrement: 0.5 g' H^-1 g is the decrease the quadratic model predicts -- MuJoCo's
         * 'improvement' test without the cancellation of a difference of two costs (which fp32 cannot resolve) */
        double dec = 0;
        for (int i = 0; i < nv; i++) dec -= grad[i] * p[i];
        if (0.5 * scale * dec < d->opt.tol) break;
        for (int r = 0; r < n; r++) jv[r] = rowdot(d, r, p);
        double pMp = 0, pMa = 0;
        for (int i = 0; i < nv; i++) {
            double s = 0;
            for (int j = 0; j < nv; j++) s += d->M[i][j] * p[j];
            Mp[i] = s; pMp += p[i] * s;
        }
        for (int i = 0; i < nv; i++) pMa += Mp[i] * acc[i];
        /* exact line search: safeguarded Newton on phi'(alpha); phi is convex and C1.  The step taken is the evaluated
         * point with the lowest cost (alpha = 0 included), so a truncated search can never increase the cost. */
        double lo = 0, hi = -1, alpha = 0, d1_0 = 0, best_alpha = 0, best_phi = 0, d1lo = 0, d1hi = 0;
        int nls = d->opt.newton_ls > 0 ? d->opt.newton_ls : 60;
        double lstol = getenv("ORA_LSTOL") ? atof(getenv("ORA_LSTOL")) : 1e-13;
        int secant = getenv("ORA_LS_SECANT") != NULL;
        for (int k = 0; k <= nls; k++) {
            double d1 = pMa + alpha * pMp, d2 = pMp, y[6], Hc[6][6];
            double phi = alpha * pMa + 0.5 * alpha * alpha * pMp;      /* Gauss part relative to alpha = 0 */
            for (int i = 0; i < n;) {
                int dim = block_dim(d, i);
                for (int q = 0; q < dim; q++) y[q] = jar[i + q] + alpha * jv[i + q];
                phi += block_cost(d, i, dim, y, ftmp + i, Hc);
                for (int q = 0; q < dim; q++) {
                    d1 -= ftmp[i + q] * jv[i + q];
                    for (int l = 0; l < dim; l++) d2 += jv[i + q] * Hc[q][l] * jv[i + l];
                }
                i += dim;
            }
            g_stat_lseval++;
            if (k == 0) { d1_0 = d1; best_phi = phi; }
            else {
                if (phi < best_phi) { best_phi = phi; best_alpha = alpha; }
                if (fabs(d1) <= lstol * fabs(d1_0) || k == nls) break;
            }
            if (d1 < 0) { lo = alpha; d1lo = d1; } else { hi = alpha; d1hi = d1; }
            double next = alpha - d1 / d2;
            if (next <= lo || (hi >= 0 && next >= hi)) {
                if (hi < 0) next = 2 * alpha + 1e-12;
                else if (secant) next = lo - d1lo * (hi - lo) / (d1hi - d1lo);
                else next = 0.5 * (lo + hi);
            }
            alpha = next;
        }
        alpha = best_alpha;
        if (getenv("ORA_NEWTON_TRACE")) fprintf(stderr, "  it %d cost %.12g |g| %.3e alpha %.4g d1_0 %.3e\n", it, cost, sqrt(gn), alpha, d1_0);
        for (int i = 0; i < nv; i++) acc[i] += alpha * p[i];
        for (int r = 0; r < n; r++) jar[r] += alpha * jv[r];
        d->solver_iters = it + 1;
        if (alpha == 0) break;
    }
    primal_eval(d, acc, jar, f, NULL);   /* forces of the final iterate */
}

static void stage_solve(ora_data *d) {
    const ora_model *m = d->m;
    int nv = m->nv, n = d->nefc;
    d->solver_iters = 0;
    memcpy(d->qacc, d->qacc_smooth, nv * sizeof(double));
    memset(d->qfrc_constraint, 0, sizeof(d->qfrc_constraint));
    if (!n) return;
    /* A = J M^-1 J^T, b = J a_smooth - aref */
    for (int r = 0; r < n; r++) {
        double *x = d->MinvJT + (size_t)r * NV_MAX;
        memcpy(x, d->efc_J[r], nv * sizeof(double));
        chol_solve(nv, d->L, x);
        d->efc_b[r] = rowdot(d, r, d->qacc_smooth) - d->efc_aref[r];
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = rowdot(d, i, d->MinvJT + (size_t)j * NV_MAX);
            A_(i, j) = A_(j, i) = s;
        }
    double *f = d->efc_force;
    /* warm start from qacc_warmstart: primal -> dual map f = -(J a - aref)/R, projected; kept only if it beats f = 0 */
    memset(f, 0, n * sizeof(double));
    /* identity key of every block: scalar rows (type, id); contacts (geom pair, ordinal within the pair) */
    int key_of[NEFC_MAX];
    {
        int prev_pair = -1, ord = 0;
        for (int r = 0; r < n; r++) {
            if (d->efc_type[r] != ROW_CONTACT) { key_of[r] = (d->efc_type[r] << 24) | d->efc_id[r]; continue; }
            const ora_contact *con = &d->con[d->efc_id[r]];
            int pair = con->geom1 | (con->geom2 << 8);
            ord = pair == prev_pair ? ord + 1 : 0;
            prev_pair = pair;
            key_of[r] = (ROW_CONTACT << 24) | pair | (ord << 16);
            for (int k = 1; k < con->dim; k++) key_of[r + k] = -1;
            r += con->dim - 1;
        }
    }
    if (d->opt.warmstart) {
        if (d->opt.warmstart == 2) {
            for (int r = 0; r < n; r++) {
                if (key_of[r] < 0) continue;
                int dim = d->efc_type[r] == ROW_CONTACT ? d->con[d->efc_id[r]].dim : 1;
                for (int c = 0; c < d->cache_n; c++)
                    if (d->cache_key[c] == key_of[r]) { for (int k = 0; k < dim; k++) f[r + k] = d->cache_f[c][k]; break; }
            }
        } else
        for (int r = 0; r < n; r++) f[r] = -(rowdot(d, r, d->qacc_warmstart) - d->efc_aref[r]) / d->efc_R[r];
        for (int r = 0; r < n; r++) {
            if (d->efc_type[r] == ROW_FLOSS) f[r] = fmin(fmax(f[r], -d->efc_floss[r]), d->efc_floss[r]);
            else if (d->efc_type[r] == ROW_LIMIT) f[r] = fmax(f[r], 0);
            else if (d->efc_type[r] == ROW_CONTACT) {
                const ora_contact *con = &d->con[d->efc_id[r]];
                if (f[r] <= 0) for (int k = 0; k < con->dim; k++) f[r + k] = 0;
                else {
                    double s = 0;
                    for (int k = 1; k < con->dim; k++) s += (f[r + k] / con->friction[k - 1]) * (f[r + k] / con->friction[k - 1]);
                    if (s > f[r] * f[r]) { double sc = f[r] / sqrt(s); for (int k = 1; k < con->dim; k++) f[r + k] *= sc; }
                }
                r += con->dim - 1;
            }
        }
        /* MuJoCo keeps its warm start only if it beats f = 0; the force cache is kept unconditionally (any feasible start
         * converges, and the comparison is a knife edge between fp32 and fp64 when both costs are close) */
        if (d->opt.warmstart != 2 && dual_cost(d, f) >= 0) memset(f, 0, n * sizeof(double));
    }
    /* block projected Gauss-Seidel */
    double trM = 0;
    for (int i = 0; i < nv; i++) trM += d->M[i][i];
    /* Sweep order.  Scalar rows first, in row order.  Contacts in the order of the CUDA solver's half-warp schedule: it
     * updates two contacts at once when they touch disjoint kinematic trees (such updates commute, so doing them one
     * after the other here is the same computation); the schedule pairs each not-yet-scheduled contact with the next
     * later one whose trees are disjoint from its own, greedily, and walks the pairs (a then b) in creation order. */
    int blk_start[NEFC_MAX], nblk = 0;
    {
        int cstart[NCON_MAX], cmask[NCON_MAX], used[NCON_MAX], nc = 0;
        for (int i = 0; i < n; i++) {
            if (d->efc_type[i] != ROW_CONTACT) { blk_start[nblk++] = i; continue; }
            const ora_contact *con = &d->con[d->efc_id[i]];
            int t1 = m->body_tree[m->geom_body[con->geom1]], t2 = m->body_tree[m->geom_body[con->geom2]];
            cstart[nc] = i; cmask[nc] = (t1 >= 0 ? 1 << t1 : 0) | (t2 >= 0 ? 1 << t2 : 0); used[nc] = 0; nc++;
            i += con->dim - 1;
        }
        for (int a = 0; a < nc; a++) {
            if (used[a]) continue;
            used[a] = 1;
            blk_start[nblk++] = cstart[a];
            for (int b = a + 1; b < nc; b++)
                if (!used[b] && !(cmask[a] & cmask[b])) { used[b] = 1; blk_start[nblk++] = cstart[b]; break; }
        }
    }
    if (d->opt.solver == 1) solve_newton(d, f);
    for (int it = 0; it < d->opt.max_iter && d->opt.solver == 0; it++) {
        double improvement = 0; /* decrease of the dual cost over this sweep (exact, block by block) */
        for (int kk = 0; kk < nblk; kk++) {
            int i = blk_start[(d->opt.sweep_mode == 1 && (it & 1)) ? nblk - 1 - kk : kk];
            int type = d->efc_type[i];
            if (type != ROW_CONTACT) {
                double res = d->efc_b[i] + d->efc_R[i] * f[i];
                for (int j = 0; j < n; j++) res += A_(i, j) * f[j];
                double old = f[i], x = old - res / (A_(i, i) + d->efc_R[i]);
                if (type == ROW_FLOSS) x = fmin(fmax(x, -d->efc_floss[i]), d->efc_floss[i]);
                else if (type == ROW_LIMIT) x = fmax(x, 0);
                f[i] = x;
                improvement -= (x - old) * (0.5 * (x - old) * (A_(i, i) + d->efc_R[i]) + res);
                continue;
            }
            const ora_contact *con = &d->con[d->efc_id[i]];
            int dim = con->dim;
            double res[6], old[6], AR[6][6];
            for (int k = 0; k < dim; k++) {
                old[k] = f[i + k];
                res[k] = d->efc_b[i + k] + d->efc_R[i + k] * f[i + k];
                for (int j = 0; j < n; j++) res[k] += A_(i + k, j) * f[j];
                for (int l = 0; l < dim; l++) AR[k][l] = A_(i + k, i + l) + (k == l ? d->efc_R[i + k] : 0);
            }
            /* (a) ray update: scale the whole block force along itself; normal-only when the block is empty */
            if (f[i] < MINVAL) {
                f[i] = fmax(0, f[i] - res[0] / AR[0][0]);
                for (int k = 1; k < dim; k++) f[i + k] = 0;
            } else {
                double vAv = 0, vr = 0;
                for (int k = 0; k < dim; k++) {
                    vr += old[k] * res[k];
                    for (int l = 0; l < dim; l++) vAv += old[k] * AR[k][l] * old[l];
                }
                if (vAv > MINVAL) {
                    double x = -vr / vAv;
                    if (old[0] + x * old[0] < 0) x = -1;
                    for (int k = 0; k < dim; k++) f[i + k] = old[k] + x * old[k];
                }
            }
            /* (b) friction rows with the normal force fixed: QCQP on the ellipsoid of radius f_n */
            if (f[i] < MINVAL) {
                for (int k = 1; k < dim; k++) f[i + k] = 0;
            } else {
                double Ac[5][5], bc[5], y[5];
                for (int k = 1; k < dim; k++) {
                    bc[k - 1] = res[k] + AR[k][0] * (f[i] - old[0]);
                    for (int l = 1; l < dim; l++) {
                        Ac[k - 1][l - 1] = AR[k][l];
                        bc[k - 1] -= AR[k][l] * old[l];
                    }
                }
                qcqp(dim - 1, Ac, bc, con->friction, f[i], y);
                for (int k = 1; k < dim; k++) f[i + k] = y[k - 1];
            }
            for (int k = 0; k < dim; k++) {
                double s = res[k];
                for (int l = 0; l < dim; l++) s += 0.5 * AR[k][l] * (f[i + l] - old[l]);
                improvement -= (f[i + k] - old[k]) * s;
            }
        }
        d->solver_iters = it + 1;
        if (improvement / trM < d->opt.tol) break;
    }
    /* noslip post-pass: friction-loss rows and contact friction rows on the unregularised A, normals fixed */
    int nos = d->opt.noslip_iter >= 0 ? d->opt.noslip_iter : m->noslip_iterations;
    for (int it = 0; it < nos; it++) {
        for (int kk = 0; kk < nblk; kk++) {
            int i = blk_start[kk];
            int type = d->efc_type[i];
            if (type == ROW_FLOSS) {
                double res = d->efc_b[i];
                for (int j = 0; j < n; j++) res += A_(i, j) * f[j];
                f[i] = fmin(fmax(f[i] - res / A_(i, i), -d->efc_floss[i]), d->efc_floss[i]);
            } else if (type == ROW_CONTACT) {
                const ora_contact *con = &d->con[d->efc_id[i]];
                int dim = con->dim;
                if (f[i] < MINVAL) {
                    for (int k = 1; k < dim; k++) f[i + k] = 0;
                } else {
                    double Ac[5][5], bc[5], y[5];
                    for (int k = 1; k < dim; k++) {
                        double res = d->efc_b[i + k];
                        for (int j = 0; j < n; j++) res += A_(i + k, j) * f[j];
                        bc[k - 1] = res;
                        for (int l = 1; l < dim; l++) {
                            Ac[k - 1][l - 1] = A_(i + k, i + l);
                            bc[k - 1] -= A_(i + k, i + l) * f[i + l];
                        }
                    }
                    qcqp(dim - 1, Ac, bc, con->friction, f[i], y);
                    for (int k = 1; k < dim; k++) f[i + k] = y[k - 1];
                }
            }
        }
    }
    if (!d->cache_frozen) d->cache_n = 0;
    for (int r = 0; r < n && !d->cache_frozen; r++) {
        if (key_of[r] < 0) continue;
        int dim = d->efc_type[r] == ROW_CONTACT ? d->con[d->efc_id[r]].dim : 1, c = d->cache_n++;
        d->cache_key[c] = key_of[r];
        for (int k = 0; k < 6; k++) d->cache_f[c][k] = k < dim ? f[r + k] : 0.0;
    }
    for (int r = 0; r < n; r++) {
        const double *x = d->MinvJT + (size_t)r * NV_MAX;
        for (int i = 0; i < nv; i++) {
            d->qacc[i] += x[i] * f[r];
            d->qfrc_constraint[i] += d->efc_J[r][i] * f[r];
        }
    }
}

/* ------------------------------------------------------------------ stage 7: Euler with implicit joint damping (mj_Euler) */
static void stage_integrate(ora_data *d) {
    const ora_model *m = d->m;
    int nv = m->nv;
    double h = m->timestep;
    static __thread double Mh[NV_MAX][NV_MAX], Lh[NV_MAX][NV_MAX];
    double acc[NV_MAX];
    memcpy(Mh, d->M, sizeof Mh);
    for (int i = 0; i < nv; i++) {
        Mh[i][i] += h * m->dof_damping[i];
        acc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    }
    chol_factor(nv, Mh, Lh);
    chol_solve(nv, Lh, acc);
    memcpy(d->qacc_warmstart, d->qacc, nv * sizeof(double));
    for (int i = 0; i < nv; i++) d->qvel[i] += h * acc[i];
    for (int j = 0; j < m->njnt; j++) {
        int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
        if (m->jnt_type[j] == JNT_FREE) {
            for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
            double w[3] = {d->qvel[da + 3], d->qvel[da + 4], d->qvel[da + 5]};
            double ang = v3normalize(w) * h;
            double s = sin(0.5 * ang), rq[4] = {cos(0.5 * ang), s * w[0], s * w[1], s * w[2]}, q[4];
            quat_mul(q, d->qpos + qa + 3, rq);
            quat_normalize(q);
            memcpy(d->qpos + qa + 3, q, sizeof q);
        } else
            d->qpos[qa] += h * d->qvel[da];
    }
}

/* ------------------------------------------------------------------ reward (reference env.py get_reward x5) */
enum { CLS_LEFT = 1, CLS_RIGHT = 2, CLS_TABLE = 4, CLS_A = 8, CLS_B = 16, CLS_PIN_A = 32, CLS_PIN_B = 64, CLS_C = 128 };
static int pair_hit(int c1, int c2, int ma, int mb) { return ((c1 & ma) && (c2 & mb)) || ((c2 & ma) && (c1 & mb)); }

static int stage_reward(ora_data *d) {
    const ora_model *m = d->m;
    int tl = 0, tr = 0, a_table = 0, b_table = 0, a_b = 0, pins = 0, a_pinb = 0;
    for (int c = 0; c < d->ncon; c++) {
        int c1 = m->geom_class[d->con[c].geom1], c2 = m->geom_class[d->con[c].geom2];
        switch (m->task_id) {
        case 0: /* InsertPeg env.py:444-461: peg=A (right hand), hole-*=B (left hand), pin */
            tr |= pair_hit(c1, c2, CLS_A, CLS_RIGHT); tl |= pair_hit(c1, c2, CLS_B, CLS_LEFT);
            a_table |= pair_hit(c1, c2, CLS_TABLE, CLS_A); b_table |= pair_hit(c1, c2, CLS_TABLE, CLS_B);
            a_b |= pair_hit(c1, c2, CLS_A, CLS_B); pins |= pair_hit(c1, c2, CLS_A, CLS_PIN_A);
            break;
        case 1: /* SlotInsertion env.py:564-578 */
            tr |= pair_hit(c1, c2, CLS_A, CLS_RIGHT); tl |= pair_hit(c1, c2, CLS_A, CLS_LEFT);
            a_table |= pair_hit(c1, c2, CLS_TABLE, CLS_A); a_b |= pair_hit(c1, c2, CLS_A, CLS_B);
            pins |= pair_hit(c1, c2, CLS_PIN_A, CLS_PIN_B);
            break;
        case 2: /* SewNeedle env.py:659-677 */
            tr |= pair_hit(c1, c2, CLS_A, CLS_RIGHT); tl |= pair_hit(c1, c2, CLS_A, CLS_LEFT);
            a_table |= pair_hit(c1, c2, CLS_TABLE, CLS_A); a_b |= pair_hit(c1, c2, CLS_A, CLS_B);
            pins |= pair_hit(c1, c2, CLS_PIN_A, CLS_PIN_B); a_pinb |= pair_hit(c1, c2, CLS_A, CLS_PIN_B);
            break;
        case 3: /* TubeTransfer env.py:756-770: tube1=A (right), tube2=B (left), ball=C, pin */
            tr |= pair_hit(c1, c2, CLS_A, CLS_RIGHT); tl |= pair_hit(c1, c2, CLS_B, CLS_LEFT);
            a_table |= pair_hit(c1, c2, CLS_TABLE, CLS_A); b_table |= pair_hit(c1, c2, CLS_TABLE, CLS_B);
            pins |= pair_hit(c1, c2, CLS_C, CLS_PIN_A);
            break;
        case 4: /* HookPackage env.py:838-852: package=A, hook=B */
            tr |= pair_hit(c1, c2, CLS_A, CLS_RIGHT); tl |= pair_hit(c1, c2, CLS_A, CLS_LEFT);
            a_table |= pair_hit(c1, c2, CLS_TABLE, CLS_A); a_b |= pair_hit(c1, c2, CLS_B, CLS_A);
            pins |= pair_hit(c1, c2, CLS_PIN_A, CLS_PIN_B);
            break;
        }
    }
    int r = 0;
    switch (m->task_id) {
    case 0:
        if (tl && tr) r = 1;
        if (tl && tr && !a_table && !b_table) r = 2;
        if (a_b && !a_table && !b_table) r = 3;
        if (pins) r = 4;
        break;
    case 1:
        if (tl && tr) r = 1;
        if (tl && tr && !a_table) r = 2;
        if (a_b && !a_table) r = 3;
        if (pins) r = 4;
        break;
    case 2:
        if (pins) d->latch = 1;
        if (tr) r = 1;
        if (tr && !a_table) r = 2;
        if (a_b && !a_table) r = 3;
        if (d->latch) r = 4;
        if (tl && !tr && !a_table && !a_pinb && d->latch) r = 5;
        break;
    case 3:
        if (tl && tr) r = 1;
        if (tl && tr && !a_table && !b_table) r = 2;
        if (pins) r = 3;
        break;
    case 4:
        if (tl && tr) r = 1;
        if (tl && tr && !a_table) r = 2;
        if (a_b && !a_table) r = 3;
        if (pins) r = 4;
        break;
    }
    d->reward = r;
    return r;
}

/* ------------------------------------------------------------------ public drivers */
static int forward_pass(ora_data *d, int frozen) {
    stage_kinematics(d);
    if (stage_inertia(d)) return -1;
    stage_collision(d);
    stage_smooth(d);
    stage_constraint_rows(d);
    d->cache_frozen = frozen;
    stage_solve(d);
    d->cache_frozen = 0;
    return 0;
}
/* physics.forward(): like mj_forward it leaves the warm-start state (qacc_warmstart / force cache) untouched */
int ora_forward(ora_data *d) { return forward_pass(d, 1); }
int ora_substep(ora_data *d) {
    if (forward_pass(d, 0)) return -1;
    stage_integrate(d);
    return 0;
}
/* trailing position pass of dm_control's Physics.step (mj_step1): contacts of the post-step configuration */
void ora_position_pass(ora_data *d) {
    stage_kinematics(d);
    stage_collision(d);
}
/* reference env.py:203-226: ctrl write (gripper un-normalise), nsub substeps, reward */
int ora_env_step(ora_data *d, const double *action, int nsub) {
    const ora_model *m = d->m;
    int nj = m->num_arms == 3 ? 21 : 14;
    for (int k = 0; k < nj; k++) {
        double a = action[k];
        if (k == 6 || k == 13) a = a * (m->act_ctrl_hi[k] - m->act_ctrl_lo[k]) + m->act_ctrl_lo[k];
        d->ctrl[k] = a;
    }
    for (int s = 0; s < nsub; s++)
        if (ora_substep(d)) return -1;
    ora_position_pass(d);
    return stage_reward(d);
}
/* reference env.py:228-249 + task reset: home pose, open fingers, ctrl = home, object poses, qvel = 0 */
void ora_reset(ora_data *d, const double *arm_pose /*21*/, const double *free_pos /*nfree*3 or NULL*/) {
    d->cache_n = 0;   /* a fresh episode has no force history */
    const ora_model *m = d->m;
    memcpy(d->qpos, m->qpos0, m->nq * sizeof(double));
    memset(d->qvel, 0, sizeof(d->qvel));
    memset(d->qacc_warmstart, 0, sizeof(d->qacc_warmstart));
    d->latch = 0;
    double open_l = m->act_ctrl_hi[6], open_r = m->act_ctrl_hi[13];
    for (int k = 0; k < 21; k++) {
        d->qpos[m->obs_qadr[k]] = arm_pose[k];
        d->ctrl[k] = arm_pose[k];
    }
    d->qpos[m->finger_qadr[0]] = d->qpos[m->finger_qadr[1]] = open_l;
    d->qpos[m->finger_qadr[2]] = d->qpos[m->finger_qadr[3]] = open_r;
    d->ctrl[6] = open_l; d->ctrl[13] = open_r;
    for (int k = 0; k < m->nfree && free_pos; k++) {
        int qa = m->free_qadr[k];
        v3cpy(d->qpos + qa, free_pos + 3 * k);
        d->qpos[qa + 3] = 1; d->qpos[qa + 4] = d->qpos[qa + 5] = d->qpos[qa + 6] = 0;
    }
    ora_position_pass(d);
    stage_reward(d);
}
void ora_agent_pos(const ora_data *d, double *out) {
    const ora_model *m = d->m;
    int nj = m->num_arms == 3 ? 21 : 14;
    for (int k = 0; k < nj; k++) {
        double q = d->qpos[m->obs_qadr[k]];
        if (k == 6 || k == 13) q = (q - m->act_ctrl_lo[k]) / (m->act_ctrl_hi[k] - m->act_ctrl_lo[k]);
        out[k] = q;
    }
}

/* ------------------------------------------------------------------ accessors for the tests (ctypes) */
int ora_nq(const ora_model *m) { return m->nq; }
int ora_nv(const ora_model *m) { return m->nv; }
int ora_nu(const ora_model *m) { return m->nu; }
int ora_nbody(const ora_model *m) { return m->nbody; }
int ora_ngeom(const ora_model *m) { return m->ngeom; }
double *ora_qpos(ora_data *d) { return d->qpos; }
double *ora_qvel(ora_data *d) { return d->qvel; }
double *ora_ctrl(ora_data *d) { return d->ctrl; }
double *ora_qacc_warmstart(ora_data *d) { return d->qacc_warmstart; }
double *ora_qacc(ora_data *d) { return d->qacc; }
double *ora_qacc_smooth(ora_data *d) { return d->qacc_smooth; }
double *ora_qfrc_bias(ora_data *d) { return d->qfrc_bias; }
double *ora_qfrc_actuator(ora_data *d) { return d->qfrc_actuator; }
double *ora_qfrc_constraint(ora_data *d) { return d->qfrc_constraint; }
double *ora_xpos(ora_data *d) { return &d->xpos[0][0]; }
double *ora_xquat(ora_data *d) { return &d->xquat[0][0]; }
double *ora_gpos(ora_data *d) { return &d->gpos[0][0]; }
double *ora_gmat(ora_data *d) { return &d->gmat[0][0]; }
double *ora_M(ora_data *d) { return &d->M[0][0]; }
int ora_M_stride(void) { return NV_MAX; }
int ora_ncon(const ora_data *d) { return d->ncon; }
int ora_nefc(const ora_data *d) { return d->nefc; }
int ora_solver_iters(const ora_data *d) { return d->solver_iters; }
int ora_reward(const ora_data *d) { return d->reward; }
/* test hook: the staged reward of an explicit contact list (geom id pairs) and latch state; clobbers the contact list */
int ora_reward_from_pairs(ora_data *d, const int *pairs, int n, int latch_in, int *latch_out) {
    if (n > NCON_MAX) return -1;
    d->ncon = n;
    for (int c = 0; c < n; c++) { d->con[c].geom1 = pairs[2 * c]; d->con[c].geom2 = pairs[2 * c + 1]; }
    d->latch = latch_in;
    int r = stage_reward(d);
    if (latch_out) *latch_out = d->latch;
    return r;
}
int *ora_latch(ora_data *d) { return &d->latch; }
double *ora_efc_force(ora_data *d) { return d->efc_force; }
double *ora_efc_aref(ora_data *d) { return d->efc_aref; }
double *ora_efc_R(ora_data *d) { return d->efc_R; }
double *ora_efc_J(ora_data *d) { return &d->efc_J[0][0]; }
/* contact c -> out[0]=dist, [1:4]=pos, [4:13]=frame, [13]=geom1, [14]=geom2, [15]=dim, [16]=excluded, [17:22]=friction */
void ora_contact_get(const ora_data *d, int c, double *out) {
    const ora_contact *k = &d->con[c];
    out[0] = k->dist;
    memcpy(out + 1, k->pos, 3 * sizeof(double));
    memcpy(out + 4, k->frame, 9 * sizeof(double));
    out[13] = k->geom1; out[14] = k->geom2; out[15] = k->dim; out[16] = k->excluded;
    memcpy(out + 17, k->friction, 5 * sizeof(double));
}
/* narrowphase of one geom pair at explicit poses (for collision unit tests); returns #contacts, rows of 13 doubles */
int ora_collide_pair(const ora_model *m, int g1, const double *pos1, const double *mat1, int g2, const double *pos2,
                     const double *mat2, int multiccd, double *out, int max_out);
