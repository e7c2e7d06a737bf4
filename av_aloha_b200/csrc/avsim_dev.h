// avsim_dev.h -- device-side model tables and batch state shared by the kernels and the host API.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   * model constants: one fp32 blob + one int32 blob per loaded model, read-only, shared by all environments
//     (hull vertices 18 014 x float4 = 288 KB dominate; everything is L2-resident after the first substep);
//   * per-environment persistent state, environment-major rows: qpos[nq] qvel[nv] ctrl[nu] warmstart[nv]
//     (+ latch, reward, status, agent_pos) -- the 1 032 B/env-step algorithmic traffic of SURVEY.md 8(d);
//   * per-environment solver scratch (contact Jacobian blocks), written and re-read inside one launch only.
#pragma once
#include <stdint.h>

#define AV_NB 32        // max bodies
#define AV_NV 41        // max dofs (TubeTransfer)
#define AV_NVP 44       // padded
#define AV_NQ 44
#define AV_NU 21
#define AV_NG 96        // max collidable geoms
#define AV_NTREE 6
#define AV_TD 8         // max dofs per kinematic tree
#define AV_MBLK (AV_NTREE * AV_TD * AV_TD)
#define AV_MTRI (AV_TD * (AV_TD + 1) / 2)   // packed lower triangle of one tree's mass block (M is symmetric; only M^-1 and L are kept full)
#define AV_MPK (AV_NTREE * AV_MTRI)
#define AV_NCON 64      // max contacts per environment (== AVSIM_MAX_CONTACTS)
#define AV_NSC 20       // max scalar constraint rows (equality + friction loss + joint limits)
#define AV_NCAND 64     // broadphase survivors per class
#define AV_MAX_ENVW 15   // environments (shared-memory slices) per block
#ifndef AV_MAX_WARPS
#define AV_MAX_WARPS 16  // warps per block of the step kernel (16 x 32 x 128 registers = the register file); AV_MAX_ENVW of them own a slice
#endif
#define AV_MIN_BLOCKS 14 // resident single-warp blocks per SM the register allocation of the forward kernel must allow
#ifndef AV_BULK_PREFETCH
#define AV_BULK_PREFETCH 0 // 1: TMA bulk prefetch of contact blocks in the solver sweep (measured slower, see avsim_solve.cuh)
#endif
#define AV_JW 16        // columns of a contact Jacobian block: 8 dofs of tree1 | 8 dofs of tree2
#define AV_NHP (AV_NV * (AV_NV + 1) / 2)   // packed lower triangle of the Newton solver's Hessian

enum { AV_JNT_FREE = 0, AV_JNT_SLIDE = 2, AV_JNT_HINGE = 3 };
enum { AV_GEOM_SPHERE = 2, AV_GEOM_CYLINDER = 5, AV_GEOM_BOX = 6, AV_GEOM_MESH = 7 };
enum { AV_PAIR_SS = 0, AV_PAIR_SB = 1, AV_PAIR_BS = 2, AV_PAIR_BB = 3, AV_PAIR_CONVEX = 4 };

struct DevModel {
    int nbody, njnt, nv, nq, ngeom, npair, nu, neq, ntree, nfree, nj_obs;
    int task_id, max_reward, num_arms, noslip_iterations, multiccd;
    float timestep, impratio, gravity[3];
    // bodies
    const int *body_parent, *body_jntadr, *body_jntnum, *body_dofadr, *body_dofnum, *body_tree, *body_lastdof;
    const int *body_treemask;   // bit k: the tree-local dof k moves this body
    const float *body_pos, *body_quat, *body_mass, *body_ipos, *body_inertia, *body_invweight0;
    const float *body_xpos0, *body_xquat0;   // world pose at qpos0 (exact for world-welded bodies)
    // trees (bodies and dofs of a tree are contiguous)
    const int *tree_bodyadr, *tree_bodynum, *tree_dofadr, *tree_dofnum;
    // joints / dofs
    const int *jnt_type, *jnt_qposadr, *jnt_dofadr, *jnt_limited;
    const float *jnt_axis, *jnt_pos, *jnt_range, *jnt_solref, *jnt_solimp;
    const int *dof_body, *dof_jnt, *dof_parent, *dof_tree, *dof_frc_limited;
    const float *dof_armature, *dof_damping, *dof_frictionloss, *dof_frc_lo, *dof_frc_hi, *dof_invweight0,
        *dof_solref, *dof_solimp;
    const float *qpos0;
    // geoms
    const int *geom_type, *geom_body, *geom_condim, *geom_hull, *geom_class, *geom_static;
    const float *geom_pos, *geom_mat, *geom_size, *geom_rbound, *geom_aabb, *geom_friction, *geom_solref,
        *geom_solimp, *geom_gap, *geom_margin;
    const float *geom_xpos0, *geom_xmat0, *geom_xaabb0;   // world pose / world AABB half extents of world-welded geoms
    const int *hull_adr, *hull_num;
    const float4 *hull_vert;
    const int *pair_geom;                    // packed g1 | g2 << 8 | type << 16
    const float *pair_rsum;                  // rbound1 + rbound2
    // equality / actuators / env tables
    const int *eq_dof1, *eq_dof2, *eq_qadr1, *eq_qadr2;
    const float *eq_polycoef, *eq_solref, *eq_solimp, *eq_invweight0;
    const int *act_dof, *act_qadr;
    const float *act_kp, *act_kv, *act_ctrl_lo, *act_ctrl_hi;
    const int *obs_qadr, *finger_qadr, *free_qadr;
    const float *reset_lo, *reset_hi;        // [nfree*3] uniform ranges of the task reset
    const int *reset_draw;                   // Philox stream layout (which draw feeds which coordinate)
    // cameras / render colours (K9)
    int ncam;
    const int *cam_body, *geom_visible, *geom_tex;
    const float *cam_pos, *cam_quat, *cam_fovy, *geom_rgba;
    const float *hull_kdop;   // [nhull][13][2] slab intervals of the 26-DOP around every hull (render stand-in for mesh geoms)
    const float *light;       // directional light: dir 3 | diffuse | headlight ambient | headlight diffuse
    const float *table_tex;   // [128][128][3] diffuse texture of the table top
    // IK tables
    const int *ik_ndof;
    const float *ik_w0, *ik_p0, *ik_site0, *ik_range;
};

struct BatchState {
    int num_envs;
    float *qpos, *qvel, *ctrl, *warm, *agent_pos;   // [B][nq|nv|nu|nv|nj]
    int *reward, *status, *latch, *ncon, *episode;  // [B]
    float *contacts;                                // [B][AV_NCON][16]
    float *qacc, *xpos, *qfrc_bias, *qacc_smooth, *mass_diag;  // debug dumps of the last forward pass
    float *scratch;                                 // [B][AV_SCRATCH_FLOATS]
    int *order;                                     // [B] environments in the order the step kernel hands them out (costliest first)
    int *queue;                                     // [1] head of that work queue
    // constraint-force cache (warm start mode 2): forces of the previous solve keyed by constraint identity
    int *fc_key, *fc_n;                             // [B][AV_NCON + AV_NSC], [B][2] (contact keys, scalar keys)
    float *fc_val;                                  // [B][AV_NCON * 6 + AV_NSC]
    int warm_mode;                                  // 1: MuJoCo-style map of qacc_warmstart, 2: force cache
    long long *env_cycles;                          // [B] SM cycles the last step kernel spent on each environment
    // split pipeline (avsim_substep_kernel / avsim_solve_kernel): per-environment head image, the solver kernel's own queue
    float *heads;                                   // [B][AV_HEAD_FLOATS]
    int *order_b, *queue_b;                         // environments sorted by the solver cycles of the previous step; queue head
    long long *env_cycles_b;                        // [B] SM cycles of the last solver launch per environment
    uint64_t seed;
    int solver_iters, noslip_iters, multiccd;
    // constraint solver: 0 = block PGS on the dual (solver_iters sweeps), 1 = Newton on the primal (the reference's solver:
    // at most newton_iters iterations, newton_ls line-search evaluations each, stop at scaled gradient < newton_tol); both
    // are followed by noslip_iters noslip sweeps
    int solver, newton_iters, newton_ls;
    float newton_tol;
    float *nw_stat;                                 // [B][4] Newton statistics of the last launch (NwStat, avsim_kernels.cuh)
    // The first `heavy_tasks` entries of the work queue hand out only `heavy_warps` environments per block (the costliest ones:
    // the queue is sorted); the block's remaining warps carry no environment and only pull pooled narrowphase items, which
    // shortens the launch's critical path (the block with the costliest environments).  0 = every task takes a full block.
    int heavy_tasks, heavy_warps;
    // warps of a block that own an environment slice (<= blockDim.y); the block's other warps are helpers without shared-memory
    // state that only pull pooled narrowphase items (registers allow 16 warps per SM, shared memory 13-15 environment slices)
    int env_warps;
    int key_pooled;   // 1: the queue's sort key (env_cycles) includes the environment's pooled narrowphase items, 0: only its own stages
    int sync;   // lockstep granularity of a block's warps: 2 = barrier after every stage, 1 = once per substep, 0 = none
};

// per-contact block in global scratch (avsim_solve.cuh): AR 21 | Lc 15 | b 6 | R 4 | mu 3 | 1/mu 3 | J[6][16] |
// geometry (pos 3, frame 9, dist 1, friction 3) written by the narrowphase, read once by the row assembly
#define AV_CBLK (52 + 6 * AV_JW + 16)
#define AV_NKEEP 32      // convex pairs per environment that survive the box filter
#define AV_CTMP 24       // floats of one pooled narrowphase result: n | normal 3 | 5 x dist | 5 x pos
#define AV_SCR_TMP (AV_NCON * AV_CBLK)
#define AV_SCRATCH_FLOATS (AV_SCR_TMP + AV_NKEEP * AV_CTMP)
