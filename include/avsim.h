/*
 * avsim.h -- C-ABI of the B200-native batched simulator behind gym_guided_vision's env.step()/reset()
 * and the diff_ik / grad_ik controllers.
 *
 * The reference has no native boundary of its own (it calls MuJoCo through dm_control and numba-JITs its IK);
 * each entry point below names the reference call it replaces so a maintainer can bind it (INTEGRATION.md
 * shows the ctypes stub).  Plain pointers and sizes only; all *_dev pointers are CUDA device pointers owned by
 * the caller (e.g. torch tensors), contiguous, environment-major.  Every function returns 0 on success or a
 * negative code; the message is available from avsim_last_error() (thread-local).  Nothing throws across the ABI.
 * Calls on one avsim_batch must be serialised by the caller; work is enqueued on the batch's stream and is
 * asynchronous with respect to the host.
 */
#ifndef AVSIM_H
#define AVSIM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct avsim_model avsim_model;
typedef struct avsim_batch avsim_batch;

#define AVSIM_OK 0
#define AVSIM_ERR_ARG -1
#define AVSIM_ERR_IO -2
#define AVSIM_ERR_CUDA -3
#define AVSIM_ERR_LIMIT -4

/* fields of avsim_get / avsim_set  (dtype, per-env length) */
enum avsim_field {
    AVSIM_QPOS = 0,        /* f32 [nq]      physics.data.qpos            (reference env.py:251-253)           */
    AVSIM_QVEL = 1,        /* f32 [nv]      physics.data.qvel                                                */
    AVSIM_CTRL = 2,        /* f32 [nu]      physics.data.ctrl            (env.py:208-215)                     */
    AVSIM_WARMSTART = 3,   /* f32 [nv]      data.qacc_warmstart                                               */
    AVSIM_AGENT_POS = 4,   /* f32 [14|21]   get_obs()['agent_pos']       (env.py:169-178)                     */
    AVSIM_REWARD = 5,      /* i32 [1]       get_reward()                 (env.py:425-472,546-589,640-690,...)  */
    AVSIM_SUCCESS = 6,     /* i32 [1]       reward == max_reward         (env.py:224)                          */
    AVSIM_NCON = 7,        /* i32 [1]       data.ncon                                                          */
    AVSIM_CONTACTS = 8,    /* f32 [AVSIM_MAX_CONTACTS][16]: dist,pos3,normal3,geom1,geom2,dim,excluded,force_n,pad4 */
    AVSIM_STATUS = 9,      /* i32 [1]       bit0 numerical blow-up, bit1 contact overflow (> 64), bit2 scalar-row overflow, bit3 block prefetch timed out, bit4 Newton Hessian not positive definite in fp32 */
    AVSIM_LATCH = 10,      /* i32 [1]       SewNeedle _threaded_needle   (env.py:602,631,673)                  */
    AVSIM_QACC = 11,       /* f32 [nv]      data.qacc of the last forward pass                                 */
    AVSIM_XPOS = 12,       /* f32 [nbody*3] data.xpos of the last forward pass                                 */
    AVSIM_QFRC_BIAS = 13,  /* f32 [nv]                                                                          */
    AVSIM_QACC_SMOOTH = 14,/* f32 [nv]                                                                          */
    AVSIM_MASS_DIAG = 15,  /* f32 [nv]      diagonal of the joint-space inertia                                */
    AVSIM_ENV_CYCLES = 16, /* i64 [1]       SM cycles the last avsim_step spent on this environment (load-balance diagnostics) */
    AVSIM_FC_KEY = 17,     /* i32 [84]      force cache: identity keys of the constraints of the last solve (64 contacts | 20 scalar rows) */
    AVSIM_FC_N = 18,       /* i32 [2]       force cache: number of contact keys, number of scalar-row keys                      */
    AVSIM_FC_VAL = 19,     /* f32 [404]     force cache: 64 x 6 contact forces | 20 scalar-row forces                           */
    AVSIM_SOLVER_STAT = 20 /* f32 [4]       Newton solver over the last launch: iterations summed over the substeps, largest scaled
                                            gradient a solve ended with, most iterations of one solve, solves that hit the cap      */
};
#define AVSIM_MAX_CONTACTS 64

/* ---- model: replaces mjcf.from_path + Physics.from_mjcf_model (reference env.py:53-56).
 * `avm_path` is a compiled model table produced by av_aloha_b200/mjcf_compile.py from the reference's MJCF. */
avsim_model *avsim_model_load(const char *avm_path, int device);
void avsim_model_free(avsim_model *m);
int avsim_model_dim(const avsim_model *m, const char *what); /* "nq","nv","nu","nbody","ngeom","njoints","max_reward","task_id","num_arms","ncam" */

/* ---- batch of environments in lockstep: replaces B x GuidedVisionEnv instances under SyncVectorEnv
 * (reference lerobot/common/envs/factory.py:50-56).  `stream` is a cudaStream_t (0 = default stream). */
avsim_batch *avsim_create(const avsim_model *m, int num_envs, uint64_t seed, void *stream);
void avsim_destroy(avsim_batch *b);

/* solver_iters: PGS sweeps per substep; noslip_iters < 0 takes the model's value (aloha_sim.xml:4);
 * multiccd < 0 takes the model's flag (aloha_sim.xml:5). */
int avsim_set_options(avsim_batch *b, int solver_iters, int noslip_iters, int multiccd);

/* constraint solver.  AVSIM_SOLVER_NEWTON (default): Newton on the primal problem -- the solver the reference runs, since
 * aloha_sim.xml:4-6 leaves MuJoCo's default -- warm-started from the previous qacc, at most max_iter iterations with an exact
 * line search of at most ls_iter evaluations, stopping when the scaled gradient drops below tol; values <= 0 keep the current
 * setting (defaults 30, 20, 3e-7: what fp32 reaches; MuJoCo's 1e-8 is an fp64 figure).  AVSIM_SOLVER_PGS: block projected Gauss-Seidel on the dual with the fixed sweep count of
 * avsim_set_options (converges to the same optimum, slowly: an approximate, fixed-cost mode).  Both are followed by the
 * model's noslip sweeps. */
#define AVSIM_SOLVER_PGS 0
#define AVSIM_SOLVER_NEWTON 1
int avsim_set_solver(avsim_batch *b, int solver, int max_iter, int ls_iter, float tol);

/* warm start of the constraint solve (PGS solver only; Newton always starts from the better of qacc_smooth and the previous qacc).  1 (default): MuJoCo's scheme, the previous qacc mapped to forces (what the reference
 * does through data.qacc_warmstart).  2: every constraint starts from the force it carried in the previous solve, matched
 * by identity (geom pair + ordinal); same optimum, but Gauss-Seidel then needs far fewer sweeps for the same accuracy
 * (profiles/r1_warmstart_accuracy.txt). */
int avsim_set_warmstart(avsim_batch *b, int mode);

/* reset: replaces GuidedVisionEnv.reset + task reset (env.py:228-249, 474-501, 513-543, 604-637, 705-735, 792-818).
 * mask_dev: u8[B] nullable (null = all envs).  free_pos_dev: f32[B][nfree][3] object positions drawn by the host
 * in the reference's np.random order, nullable (null = Philox draw on the device from (seed, env, episode)). */
int avsim_reset(avsim_batch *b, const uint8_t *mask_dev, const float *free_pos_dev);

/* step: replaces GuidedVisionEnv.step / step_action (env.py:203-226, 255-269): ctrl write with gripper
 * un-normalisation, nsubsteps x mj_step, trailing position pass, reward, agent_pos.  action_dev: f32[B][14|21]. */
int avsim_step(avsim_batch *b, const float *action_dev, int nsubsteps);

/* forward: replaces physics.forward() (env.py:244,253): position + velocity + acceleration stages, no integration. */
int avsim_forward(avsim_batch *b);

int avsim_get(avsim_batch *b, int field, void *dst_dev);
int avsim_set(avsim_batch *b, int field, const void *src_dev);

/* render: replaces physics.render(height, width, camera_id) for each configured camera (reference env.py:180-188, 195-200).
 * cam_ids_host: indices into the model's camera list (names in the model's .json sidecar), dst_dev: u8 [B][ncam][H][W][3].
 * Ray-cast of the physics geoms (mesh geoms as the 26-DOP of their hulls, scene lights, table texture); W must be a multiple of 4. */
int avsim_render(avsim_batch *b, const int *cam_ids_host, int ncam, int H, int W, uint8_t *dst_dev);
/* test hook: the same primary rays, every pixel = index of the geom it hits (all three channels; 255 = background) */
int avsim_render_ids(avsim_batch *b, const int *cam_ids_host, int ncam, int H, int W, uint8_t *dst_dev);

/* device-resident observation path: replaces the image half of lerobot's preprocess_observation (reference
 * lerobot/lerobot/common/envs/utils.py:37-50: channel-last u8 -> channel-first f32 in [0,1], fp32 division by 255,
 * bit-exact).  src_dev: u8 [n_images][H][W][3] (what avsim_render wrote, any leading [B][ncam] flattened),
 * dst_dev: f32 [n_images][3][H][W].  Stand-alone: needs no batch; runs on `stream` of `device`. */
int avsim_pixels_to_float(const uint8_t *src_dev, int64_t n_images, int H, int W, float *dst_dev, int device, void *stream);

/* host-buffer convenience path used by the gym-facing wrapper (e2e metric): copies happen inside the call.  status_host
 * (nullable): the AVSIM_STATUS word of every environment after the step -- the wrapper auto-resets environments whose bit 0
 * (numerical blow-up) is set, the analogue of MuJoCo's mj_checkAcc warning + reset. */
int avsim_step_host(avsim_batch *b, const float *action_host, int nsubsteps, float *agent_pos_host, int32_t *reward_host,
                    int32_t *status_host);

/* number of kernel launches issued by this batch so far (bench.py's gpu_launches) */
int64_t avsim_launch_count(const avsim_batch *b);
/* how this batch is launched (chosen by avsim_create from the batch size, or by the AVSIM_* diagnostic variables):
 * out[0] = 1 split pipeline (substep + solve kernels) / 0 fused step kernel, out[1] = environment groups (streams),
 * out[2] = warps per block, out[3] = of which own an environment slice, out[4] = warps per solve block, out[5] = SMs.
 * Diagnostics only: no reference call corresponds (the reference steps one environment per process). */
int avsim_launch_shape(const avsim_batch *b, int out[6]);

/* diagnostics: per-stage SM cycle counters summed over warps (only in a -DAVSIM_PROFILE build; otherwise returns
 * AVSIM_ERR_ARG).  Stages: load, kinematics, inertia, broadphase, primitive narrowphase, convex narrowphase, smooth,
 * scalar rows, contact rows, solve, integrate, outputs.  Returns the number of stages. */
int avsim_stage_cycles(uint64_t *out_host, int n, int reset);

/* ---- IK controllers (reference data_collection_scripts/diff_ik.py:51-90, grad_ik.py:8-99).
 * arm: 0 left, 1 right, 2 middle.  q/pos/quat_wxyz/q_out are device f32 arrays of n rows. */
typedef struct {
    float k_pos, k_ori, damping, max_angvel, integration_dt;
    float k_null[7], q0[7];
    int iterations;
} avsim_diffik_params;
typedef struct {
    float step_size, min_cost_delta, position_weight, rotation_weight, position_threshold, rotation_threshold,
        max_pos_diff, max_rot_diff, joint_p;
    float joint_center_weight[7], joint_displacement_weight[7];
    int max_iterations;
} avsim_gradik_params;
int avsim_diffik(const avsim_model *m, int arm, const float *q_dev, const float *pos_dev, const float *quat_wxyz_dev,
                 int n, const avsim_diffik_params *p, float *q_out_dev, void *stream);
int avsim_gradik(const avsim_model *m, int arm, const float *q_dev, const float *pos_dev, const float *quat_wxyz_dev,
                 int n, const avsim_gradik_params *p, float *q_out_dev, void *stream);
/* product-of-exponentials forward kinematics of an arm's end-effector site (kinematics.py:7-26): out f32[n][16] */
int avsim_fk(const avsim_model *m, int arm, const float *q_dev, int n, float *T_out_dev, void *stream);
/* space Jacobian of the same site, rows re-ordered to [v; w] (kinematics.py:28-52): out f32[n][6][ndof] (ndof = 6, 6, 7) */
int avsim_jac(const avsim_model *m, int arm, const float *q_dev, int n, float *J_out_dev, void *stream);

/* ---- SO(3) / SE(3) helpers of the reference's transform_utils.py (data_collection_scripts/transform_utils.py), batched:
 * n items, fp64 in and out (the reference computes in float64), quaternions (x, y, z, w) like the reference's internals.
 *   AVSIM_XF_MAT2QUAT          a[9]            -> out[4]    mat2quat 9-49 (w >= 0)
 *   AVSIM_XF_QUAT2MAT          a[4]            -> out[9]    quat2mat 52-79, including its float32 round trip
 *   AVSIM_XF_QUAT2AXISANGLE    a[4]            -> out[3]    quat2axisangle 82-106
 *   AVSIM_XF_AXISANGLE2QUAT    a[3]            -> out[4]    axisangle2quat 108-133
 *   AVSIM_XF_ANGULAR_ERROR     a[9] desired, b[9] current -> out[3]   angular_error 183-194
 *   AVSIM_XF_LIMIT_POSE        a[12] current (pos | mat), b[12] target, p0 = max_pos_diff, p1 = max_rot_diff -> out[12]   limit_pose 263-287
 *   AVSIM_XF_EXP2MAT           a[7] (w | v | theta)        -> out[16]  exp2mat 239-261
 *   AVSIM_XF_ADJOINT           a[16]           -> out[36]   adjoint 289-301
 *   AVSIM_XF_WITHIN_POSE       a[12], b[12], p0 / p1 = position / rotation threshold -> out[1] (1.0 / 0.0)   within_pose_threshold 196-201 */
enum avsim_xf_op { AVSIM_XF_MAT2QUAT = 0, AVSIM_XF_QUAT2MAT, AVSIM_XF_QUAT2AXISANGLE, AVSIM_XF_AXISANGLE2QUAT, AVSIM_XF_ANGULAR_ERROR,
                   AVSIM_XF_LIMIT_POSE, AVSIM_XF_EXP2MAT, AVSIM_XF_ADJOINT, AVSIM_XF_WITHIN_POSE };
int avsim_transform(int op, const double *a_dev, const double *b_dev, int n, double p0, double p1, double *out_dev, int device,
                    void *stream);

const char *avsim_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* AVSIM_H */
