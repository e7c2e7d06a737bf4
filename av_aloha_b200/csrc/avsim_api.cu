// avsim_api.cu -- C-ABI of libavsim.so (include/avsim.h): model upload, batch state, kernel launches.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/avsim.h"
#include "avsim_ik.cuh"
#include "avsim_kernels.cuh"
#include "avsim_model_pack.h"
#include "avsim_obs.cuh"
#include "avsim_render.cuh"

#define AV_SORT_MAX 8192
#define AV_SPLIT_MIN_ENVS_PER_SM 17
#define AV_GRADIK_GROUP_MAX 4096   // crossover measured between 4096 and 16384 problems (profiles/r2_ik_bench.txt)
#define AV_MAX_GROUPS 8
#ifndef AV_DEFAULT_HEAVY_TASKS
#define AV_DEFAULT_HEAVY_TASKS 0
#define AV_DEFAULT_HEAVY_WARPS 0
#endif
#ifndef AV_DEFAULT_WARPS
#define AV_DEFAULT_WARPS 16   // 16 warps x 128 registers fill the register file of an SM
#endif
#ifndef AV_DEFAULT_ENVW
#define AV_DEFAULT_ENVW 11   // 11 slices of 11.8 KB keep the shared-memory carve-out at 132 KB (~124 KB of L1); measured best (profiles/r1_sweeps.txt)
#endif

// The default slice count is chosen for the shared-memory / L1 split it leaves: growing EnvS past
// this bound silently moves the kernel to the next carve-out step (132 -> 164 KB) and costs ~6 % -- fail the build instead.
static_assert(AV_DEFAULT_ENVW * sizeof(EnvS) + 512 /* static shared memory */ + 1024 /* per-block reserve */ <= 132 * 1024,
              "EnvS grew: AV_DEFAULT_ENVW slices no longer fit the 132 KB shared-memory carve-out");

static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, const char *a = "", const char *b = "") {
    snprintf(g_err, sizeof g_err, fmt, a, b);
    return code;
}
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) return fail(AVSIM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define CUP(call)                                                                             \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) { fail(AVSIM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); return nullptr; } \
    } while (0)

// ------------------------------------------------------------------ model
struct avsim_model {
    DevModel dm;
    avpack::PackedModel pk;
    int device;
    float *fblob = nullptr;
    int *iblob = nullptr;
    float *hull = nullptr;
    float *home_dev = nullptr;
};

extern "C" const char *avsim_last_error(void) { return g_err; }

extern "C" avsim_model *avsim_model_load(const char *avm_path, int device) {
    avsim_model *M = new avsim_model();
    if (!avpack::pack_model(avm_path, M->pk)) {
        fail(AVSIM_ERR_IO, "avsim_model_load('%s'): %s", avm_path ? avm_path : "(null)", M->pk.error.c_str());
        delete M;
        return nullptr;
    }
    CUP(cudaSetDevice(device));
    M->device = device;
    avpack::Packer &P = M->pk.P;
    CUP(cudaMalloc(&M->fblob, P.fdata.size() * sizeof(float)));
    CUP(cudaMalloc(&M->iblob, P.idata.size() * sizeof(int)));
    CUP(cudaMalloc(&M->hull, M->pk.hull4.size() * sizeof(float)));
    CUP(cudaMemcpy(M->fblob, P.fdata.data(), P.fdata.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUP(cudaMemcpy(M->iblob, P.idata.data(), P.idata.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUP(cudaMemcpy(M->hull, M->pk.hull4.data(), M->pk.hull4.size() * sizeof(float), cudaMemcpyHostToDevice));
    M->pk.relocate(M->fblob, M->iblob, M->hull);
    M->dm = M->pk.dm;
    CUP(cudaMalloc(&M->home_dev, sizeof AV_HOME));
    CUP(cudaMemcpy(M->home_dev, AV_HOME, sizeof AV_HOME, cudaMemcpyHostToDevice));
    return M;
}

extern "C" void avsim_model_free(avsim_model *m) {
    if (!m) return;
    cudaFree(m->fblob); cudaFree(m->iblob); cudaFree(m->hull); cudaFree(m->home_dev);
    delete m;
}

extern "C" int avsim_model_dim(const avsim_model *m, const char *what) {
    if (!m || !what) return fail(AVSIM_ERR_ARG, "avsim_model_dim: null argument");
    const DevModel &d = m->dm;
    std::string w(what);
    if (w == "nq") return d.nq;
    if (w == "nv") return d.nv;
    if (w == "nu") return d.nu;
    if (w == "nbody") return d.nbody;
    if (w == "ngeom") return d.ngeom;
    if (w == "njoints") return d.nj_obs;
    if (w == "nfree") return d.nfree;
    if (w == "max_reward") return d.max_reward;
    if (w == "task_id") return d.task_id;
    if (w == "num_arms") return d.num_arms;
    if (w == "ncam") return d.ncam;
    return fail(AVSIM_ERR_ARG, "avsim_model_dim: unknown dimension '%s'", what);
}

// ------------------------------------------------------------------ batch
struct avsim_batch {
    const avsim_model *model;
    BatchState st;
    cudaStream_t stream;
    int grid, fwd_grid, warps, envw, solve_grid, split;
    // split pipeline: the environments are cut into `ngroups` contiguous groups, each stepped by its own launch chain on its own
    // stream, so that the tail of one group's solver kernel (a few environments that need many Newton iterations) overlaps with
    // the other groups' kernels instead of idling the GPU
    int ngroups = 1, per_sm = 1, sms = 1, solve_per_sm = 1, solve_warps = AV_MAX_WARPS;
    cudaStream_t gstream[AV_MAX_GROUPS] = {};
    cudaEvent_t ev_start = nullptr, ev_done[AV_MAX_GROUPS] = {};
    int64_t launches = 0;
    std::vector<void *> allocs;
    float *h_action = nullptr, *h_agent = nullptr;   // pinned staging for the host-buffer path
    int32_t *h_reward = nullptr, *h_status = nullptr;
    float *d_action = nullptr;
    float *d_rpose = nullptr;   // render: world pose of every geom and camera, [B][ngeom + ncam][12]
    float4 *d_rrect = nullptr;  // render: screen 8-DOP + nearest depth of every geom per requested camera, [B][8][ngeom][3]
    int *d_camids = nullptr;
};

template <typename T>
static bool dalloc(avsim_batch *b, T **p, size_t n) {
    if (cudaMalloc(p, n * sizeof(T)) != cudaSuccess) return false;
    cudaMemset(*p, 0, n * sizeof(T));
    b->allocs.push_back(*p);
    return true;
}

extern "C" avsim_batch *avsim_create(const avsim_model *m, int num_envs, uint64_t seed, void *stream) {
    if (!m || num_envs <= 0) { fail(AVSIM_ERR_ARG, "avsim_create: bad arguments"); return nullptr; }
    CUP(cudaSetDevice(m->device));
    avsim_batch *b = new avsim_batch();
    b->model = m;
    b->stream = (cudaStream_t)stream;
    const DevModel &d = m->dm;
    BatchState &s = b->st;
    memset(&s, 0, sizeof s);
    s.num_envs = num_envs; s.seed = seed;
    s.solver_iters = 20; s.noslip_iters = d.noslip_iterations; s.multiccd = d.multiccd; s.warm_mode = 1;
    s.solver = AVSIM_SOLVER_NEWTON; s.newton_iters = 30; s.newton_ls = 20; s.newton_tol = 3e-7f;   // the reference's solver (aloha_sim.xml:4-6); tolerance: see DESIGN.md section 4
    if (const char *et = getenv("AVSIM_NEWTON_TOL")) s.newton_tol = (float)atof(et);   // diagnostic override
    size_t B = num_envs;
    bool ok = dalloc(b, &s.qpos, B * d.nq) && dalloc(b, &s.qvel, B * d.nv) && dalloc(b, &s.ctrl, B * d.nu) &&
              dalloc(b, &s.warm, B * d.nv) && dalloc(b, &s.agent_pos, B * d.nj_obs) && dalloc(b, &s.reward, B) &&
              dalloc(b, &s.status, B) && dalloc(b, &s.latch, B) && dalloc(b, &s.ncon, B) && dalloc(b, &s.episode, B) &&
              dalloc(b, &s.contacts, B * AV_NCON * 16) && dalloc(b, &s.qacc, B * d.nv) && dalloc(b, &s.xpos, B * 3 * d.nbody) &&
              dalloc(b, &s.qfrc_bias, B * d.nv) && dalloc(b, &s.qacc_smooth, B * d.nv) && dalloc(b, &s.mass_diag, B * d.nv) &&
              dalloc(b, &s.scratch, B * AV_SCRATCH_FLOATS) && dalloc(b, &s.env_cycles, B) && dalloc(b, &s.fc_key, B * (AV_NCON + AV_NSC)) && dalloc(b, &s.fc_n, B * 2) &&
              dalloc(b, &s.fc_val, B * (AV_NCON * 6 + AV_NSC)) && dalloc(b, &s.nw_stat, B * 4) && dalloc(b, &s.heads, B * AV_HEADX_FLOATS) && dalloc(b, &s.order_b, B) && dalloc(b, &s.queue_b, AV_MAX_GROUPS) && dalloc(b, &s.env_cycles_b, B) && dalloc(b, &s.order, B) && dalloc(b, &s.queue, AV_MAX_GROUPS) && dalloc(b, &b->d_action, B * d.nj_obs);
    if (!ok) { fail(AVSIM_ERR_CUDA, "avsim_create: device allocation failed"); avsim_destroy(b); return nullptr; }
    if (cudaMallocHost(&b->h_action, B * d.nj_obs * sizeof(float)) != cudaSuccess ||
        cudaMallocHost(&b->h_agent, B * d.nj_obs * sizeof(float)) != cudaSuccess ||
        cudaMallocHost(&b->h_reward, B * sizeof(int32_t)) != cudaSuccess || cudaMallocHost(&b->h_status, B * sizeof(int32_t)) != cudaSuccess) {
        fail(AVSIM_ERR_CUDA, "avsim_create: pinned allocation failed");
        avsim_destroy(b);
        return nullptr;
    }
    // step kernel: persistent blocks of W warps (W environments in lockstep, see avsim_kernels.cuh); W x sizeof(EnvS)
    // of dynamic shared memory.  Diagnostic overrides: AVSIM_WARPS (warps per block), AVSIM_BLOCKS (blocks per SM).
    const char *ew = getenv("AVSIM_WARPS"), *eb = getenv("AVSIM_BLOCKS"), *es = getenv("AVSIM_SYNC");
    s.sync = es ? atoi(es) : 2;   // barrier after every stage; the sweeps free-run (measured: 39.8 vs 40.5 ms with per-sweep barriers)
    int sms = 0, smem_sm = 0, smem_blk = 0;
    CUP(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device));
    CUP(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, m->device));
    CUP(cudaDeviceGetAttribute(&smem_blk, cudaDevAttrMaxSharedMemoryPerBlockOptin, m->device));
    int esz = (int)sizeof(EnvS);
    // block = `warps` warps, of which the first `envw` own an environment slice in shared memory (AVSIM_WARPS / AVSIM_ENVW)
    const char *ee = getenv("AVSIM_ENVW");
    b->warps = ew ? atoi(ew) : AV_DEFAULT_WARPS;
    b->warps = std::max(1, std::min(b->warps, AV_MAX_WARPS));
    // environment groups of the split pipeline (decided first: the block shapes below follow the size of a group)
    const char *eg = getenv("AVSIM_GROUPS");
    int ng = eg ? atoi(eg) : 3;
    ng = std::max(1, std::min(ng, AV_MAX_GROUPS));
    while (ng > 1 && num_envs / ng < AV_DEFAULT_ENVW * sms / 2) ng--;   // a group should still be a good fraction of a wave
    // split pipeline from ~17 environments per SM on; below, the fused kernel (one launch per env.step: no launch gaps, no record
    // round trips) is faster: B = 128 / 512 / 1024 / 2048: 16.0 / 19.1 / 24.2 / 27.1 ms fused vs 17.1 / 20.5 / 25.8 / 29.2 split;
    // B = 4096: 47.5 fused vs 40.0 split (profiles/r2_sweeps.txt)
    const char *esp = getenv("AVSIM_SPLIT");
    b->split = esp ? atoi(esp) : (num_envs > AV_SPLIT_MIN_ENVS_PER_SM * sms ? 1 : 0);
    // Block shapes follow the batch size.  The step is bound by the serial latency of an environment, and an environment with
    // fewer lockstep partners (more helper warps on its hull pairs, more L1, no waiting for a partner's Newton iterations) runs
    // a substep faster: alone on an SM 0.8 ms instead of 1.25 ms.  Fused kernel: ONE round of blocks over the SMs while that
    // needs at most 7 slices per block (ceil(B / SMs)), above that about two rounds (ceil(B / 2 SMs) slices).  Two rounds of small
    // blocks win when environments differ in cost (heaviest first, the second round is the cheap ones: bench workload, B = 1024:
    // 4 slices 21.1 ms vs 7 slices 24.2; B = 1536: 6 slices 23.2 vs 11 slices 30.8) and lose badly when they do not (policy
    // rollout around the home pose, B = 1024: 11.8 vs 7.8 ms; B = 512: 2 slices 10.5 vs 4 slices 6.2) -- profiles/r2_sweeps.txt.
    // Split pipeline: one round per group, ceil(B / groups / SMs) slices, 11 from 9 up (9, 10 measured slower at 4096).
    const int per_sm_envs = std::max(1, ((num_envs + ng - 1) / ng + sms - 1) / sms);
    const int one_round = std::max(1, (num_envs + sms - 1) / sms), two_rounds = std::max(1, (num_envs + 2 * sms - 1) / (2 * sms));
    const int fused_envs = one_round <= 7 ? one_round : two_rounds;
    const int auto_envw = b->split ? (per_sm_envs >= 9 ? AV_DEFAULT_ENVW : per_sm_envs) : std::min(AV_DEFAULT_ENVW, fused_envs);
    b->envw = ee ? atoi(ee) : std::min(b->warps, auto_envw);
    b->envw = std::max(1, std::min(b->envw, std::min(b->warps, std::min(AV_MAX_ENVW, (smem_blk - 64) / esz))));
    s.env_warps = b->envw;
    const char *ek = getenv("AVSIM_KEY");
    s.key_pooled = ek ? atoi(ek) : 1;
    int per_sm = eb ? atoi(eb) : std::max(1, smem_sm / (b->envw * esz + 1024));
    per_sm = std::max(1, std::min(per_sm, smem_sm / (b->envw * esz + 1024)));
    per_sm = std::max(1, std::min(per_sm, 64 / b->warps));                    // 64 resident warps per SM
    // the attribute belongs to the function, not to the batch: batches of different shapes coexist in one process, so it is
    // always set to the largest block any of them may launch
    const int max_slices = std::min(AV_MAX_ENVW, (smem_blk - 64) / esz);
    CUP(cudaFuncSetAttribute(avsim_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_slices * esz));
    CUP(cudaFuncSetAttribute(avsim_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, esz));
    CUP(cudaFuncSetAttribute(avsim_substep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_slices * esz));
    {   // solver kernel: one block per SM of `solve_warps` phase-locked warps, one environment slice each
        const char *esw = getenv("AVSIM_SOLVE_WARPS");
        int sw = esw ? atoi(esw) : std::min(8, per_sm_envs);   // 8 slices leave the SM ~170 KB of L1 for the contact blocks; 6..16 measure within 2 % (profiles/r2_sweeps.txt)
        sw = std::max(1, std::min(sw, std::min(AV_MAX_WARPS, (smem_blk - 1024) / (int)AV_SOLVER_SLICE_BYTES)));
        b->solve_warps = sw;
        CUP(cudaFuncSetAttribute(avsim_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 std::min(AV_MAX_WARPS, (smem_blk - 1024) / (int)AV_SOLVER_SLICE_BYTES) * (int)AV_SOLVER_SLICE_BYTES));
        b->solve_grid = std::min((num_envs + sw - 1) / sw, sms);
        b->solve_per_sm = 1;
        b->ngroups = ng;
        for (int g = 0; g < ng && ng > 1; g++) {
            CUP(cudaStreamCreateWithFlags(&b->gstream[g], cudaStreamNonBlocking));
            CUP(cudaEventCreateWithFlags(&b->ev_done[g], cudaEventDisableTiming));
        }
        if (ng > 1) CUP(cudaEventCreateWithFlags(&b->ev_start, cudaEventDisableTiming));
    }
    b->sms = sms;
    CUP(cudaFuncSetAttribute(avsim_render_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, esz));
    CUP(cudaFuncSetAttribute(avsim_order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AV_SORT_MAX * (int)sizeof(unsigned long long)));
    // heavy tasks (BatchState::heavy_tasks): only when the batch is at least two rounds of resident slots, so that the slots the
    // helper warps give up are a small share of the launch.  Diagnostic override: AVSIM_HEAVY="tasks:warps".
    const char *eh = getenv("AVSIM_HEAVY");
    int ht = AV_DEFAULT_HEAVY_TASKS, hw = AV_DEFAULT_HEAVY_WARPS;
    if (eh && sscanf(eh, "%d:%d", &ht, &hw) != 2) { ht = 0; hw = 0; }
    if (ht < 0 || hw < 1 || hw >= b->envw || num_envs < 2 * per_sm * sms * b->envw || (long long)ht * hw > num_envs / 4) ht = 0;
    s.heavy_tasks = ht; s.heavy_warps = ht ? hw : 0;
    int tasks = ht + (num_envs - ht * s.heavy_warps + b->envw - 1) / b->envw;
    b->grid = std::min(tasks, per_sm * sms);
    b->per_sm = per_sm;
    b->fwd_grid = std::min(num_envs, 8 * sms);
    if (avsim_reset(b, nullptr, nullptr) != 0) { avsim_destroy(b); return nullptr; }
    return b;
}

extern "C" void avsim_destroy(avsim_batch *b) {
    if (!b) return;
    cudaStreamSynchronize(b->stream);
    for (int g = 0; g < AV_MAX_GROUPS; g++) {
        if (b->gstream[g]) { cudaStreamSynchronize(b->gstream[g]); cudaStreamDestroy(b->gstream[g]); }
        if (b->ev_done[g]) cudaEventDestroy(b->ev_done[g]);
    }
    if (b->ev_start) cudaEventDestroy(b->ev_start);
    for (void *p : b->allocs) cudaFree(p);
    if (b->h_action) cudaFreeHost(b->h_action);
    if (b->h_agent) cudaFreeHost(b->h_agent);
    if (b->h_reward) cudaFreeHost(b->h_reward);
    if (b->h_status) cudaFreeHost(b->h_status);
    delete b;
}

extern "C" int avsim_set_options(avsim_batch *b, int solver_iters, int noslip_iters, int multiccd) {
    if (!b || solver_iters < 0) return fail(AVSIM_ERR_ARG, "avsim_set_options: bad arguments");
    b->st.solver_iters = solver_iters;
    b->st.noslip_iters = noslip_iters >= 0 ? noslip_iters : b->model->dm.noslip_iterations;
    b->st.multiccd = multiccd >= 0 ? multiccd : b->model->dm.multiccd;
    return AVSIM_OK;
}

extern "C" int avsim_set_solver(avsim_batch *b, int solver, int max_iter, int ls_iter, float tol) {
    if (!b || (solver != AVSIM_SOLVER_PGS && solver != AVSIM_SOLVER_NEWTON)) return fail(AVSIM_ERR_ARG, "avsim_set_solver: solver must be AVSIM_SOLVER_PGS or AVSIM_SOLVER_NEWTON");
    b->st.solver = solver;
    if (max_iter > 0) b->st.newton_iters = max_iter;
    if (ls_iter > 0) b->st.newton_ls = ls_iter;
    if (tol > 0.f) b->st.newton_tol = tol;
    return AVSIM_OK;
}

extern "C" int avsim_set_warmstart(avsim_batch *b, int mode) {
    if (!b || (mode != 1 && mode != 2)) return fail(AVSIM_ERR_ARG, "avsim_set_warmstart: mode must be 1 (qacc map) or 2 (force cache)");
    b->st.warm_mode = mode;
    return AVSIM_OK;
}

static int launch_forward(avsim_batch *b, const uint8_t *mask_dev, int publish_reward) {
    avsim_forward_kernel<<<b->fwd_grid, 32, sizeof(EnvS), b->stream>>>(b->model->dm, b->st, mask_dev, publish_reward);
    b->launches++;
    CU(cudaGetLastError());
    return AVSIM_OK;
}

extern "C" int avsim_reset(avsim_batch *b, const uint8_t *mask_dev, const float *free_pos_dev) {
    if (!b) return fail(AVSIM_ERR_ARG, "avsim_reset: null batch");
    CU(cudaSetDevice(b->model->device));
    int n = b->st.num_envs;
    avsim_reset_kernel<<<(n + 127) / 128, 128, 0, b->stream>>>(b->model->dm, b->st, mask_dev, free_pos_dev, b->model->home_dev);
    b->launches++;
    CU(cudaGetLastError());
    return launch_forward(b, mask_dev, 0);   // physics.forward() + first observation of the reset envs (reference env.py:244-246)
}

static void launch_order(avsim_batch *b, const BatchState &st, cudaStream_t stream) {
    // queue order: costliest environment of the previous step first; one sorting block per 8192 environments, chunks interleaved
    int n = st.num_envs, n2 = 1, nch = (n + AV_SORT_MAX - 1) / AV_SORT_MAX;
    while (n2 < std::min(n, AV_SORT_MAX)) n2 <<= 1;
    avsim_order_kernel<<<nch, 1024, n2 * sizeof(unsigned long long), stream>>>(st, n2, AV_SORT_MAX);
    b->launches++;
}

// the view of environments [e0, e0 + n) of a batch: every per-environment array offset, own queue heads (slot g)
static BatchState group_view(const avsim_batch *b, int e0, int n, int g) {
    const DevModel &d = b->model->dm;
    BatchState v = b->st;
    size_t o = (size_t)e0;
    v.num_envs = n;
    v.qpos += o * d.nq; v.qvel += o * d.nv; v.ctrl += o * d.nu; v.warm += o * d.nv; v.agent_pos += o * d.nj_obs;
    v.reward += o; v.status += o; v.latch += o; v.ncon += o; v.episode += o;
    v.contacts += o * AV_NCON * 16;
    v.qacc += o * d.nv; v.xpos += o * 3 * d.nbody; v.qfrc_bias += o * d.nv; v.qacc_smooth += o * d.nv; v.mass_diag += o * d.nv;
    v.scratch += o * AV_SCRATCH_FLOATS;
    v.order += o; v.queue += g; v.order_b += o; v.queue_b += g;
    v.fc_key += o * (AV_NCON + AV_NSC); v.fc_n += o * 2; v.fc_val += o * (AV_NCON * 6 + AV_NSC);
    v.env_cycles += o; v.env_cycles_b += o; v.heads += o * AV_HEADX_FLOATS; v.nw_stat += o * 4;
    return v;
}

extern "C" int avsim_step(avsim_batch *b, const float *action_dev, int nsubsteps) {
    if (!b || nsubsteps < 0) return fail(AVSIM_ERR_ARG, "avsim_step: bad arguments");
    CU(cudaSetDevice(b->model->device));
    if (b->split && b->st.solver == AVSIM_SOLVER_NEWTON) {
        // split pipeline (avsim_kernels.cuh): lockstep substep kernel + free-running solver kernel, 2 nsub + 1 launches per group
        const DevModel &d = b->model->dm;
        int G = b->ngroups, n = b->st.num_envs;
        if (G > 1) CU(cudaEventRecord(b->ev_start, b->stream));
        for (int g = 0; g < G; g++) {
            cudaStream_t st = G > 1 ? b->gstream[g] : b->stream;
            int e0 = (int)((long long)n * g / G), e1 = (int)((long long)n * (g + 1) / G), ng = e1 - e0;
            if (G > 1) CU(cudaStreamWaitEvent(st, b->ev_start, 0));
            BatchState v = group_view(b, e0, ng, g), vb = v;   // vb: the solver kernel's queue, same sort keyed by its own cycles
            vb.env_cycles = v.env_cycles_b; vb.order = v.order_b; vb.queue = v.queue_b;
            launch_order(b, v, st);
            launch_order(b, vb, st);
            int tasks = (ng + b->envw - 1) / b->envw;
            int grid = std::min(tasks, b->per_sm * b->sms), sgrid = std::min((ng + b->solve_warps - 1) / b->solve_warps, b->sms);
            const float *act = action_dev ? action_dev + (size_t)e0 * d.nj_obs : nullptr;
            for (int s = 0; s <= nsubsteps; s++) {
                avsim_substep_kernel<<<grid, dim3(32, b->warps), sizeof(EnvS) * b->envw, st>>>(d, v, act, s, nsubsteps);
                if (s < nsubsteps) avsim_solve_kernel<<<sgrid, dim3(32, b->solve_warps), b->solve_warps * AV_SOLVER_SLICE_BYTES, st>>>(d, v);
            }
            b->launches += 2 * nsubsteps + 1;
            if (G > 1) {
                CU(cudaEventRecord(b->ev_done[g], st));
                CU(cudaStreamWaitEvent(b->stream, b->ev_done[g], 0));
            }
        }
    } else {
        launch_order(b, b->st, b->stream);
        avsim_step_kernel<<<b->grid, dim3(32, b->warps), sizeof(EnvS) * b->envw, b->stream>>>(b->model->dm, b->st, action_dev, nsubsteps);
        b->launches++;
    }
    CU(cudaGetLastError());
    return AVSIM_OK;
}

extern "C" int avsim_forward(avsim_batch *b) {
    if (!b) return fail(AVSIM_ERR_ARG, "avsim_forward: null batch");
    CU(cudaSetDevice(b->model->device));
    return launch_forward(b, nullptr, 1);   // set_qpos(); get_reward() reads the reward of the forward state (reference env.py:251-253, 546-589)
}

static int field_ptr(avsim_batch *b, int field, void **p, size_t *bytes) {
    const DevModel &d = b->model->dm;
    BatchState &s = b->st;
    size_t B = s.num_envs;
    switch (field) {
    case AVSIM_QPOS: *p = s.qpos; *bytes = B * d.nq * 4; break;
    case AVSIM_QVEL: *p = s.qvel; *bytes = B * d.nv * 4; break;
    case AVSIM_CTRL: *p = s.ctrl; *bytes = B * d.nu * 4; break;
    case AVSIM_WARMSTART: *p = s.warm; *bytes = B * d.nv * 4; break;
    case AVSIM_AGENT_POS: *p = s.agent_pos; *bytes = B * d.nj_obs * 4; break;
    case AVSIM_REWARD: *p = s.reward; *bytes = B * 4; break;
    case AVSIM_NCON: *p = s.ncon; *bytes = B * 4; break;
    case AVSIM_CONTACTS: *p = s.contacts; *bytes = B * AV_NCON * 16 * 4; break;
    case AVSIM_STATUS: *p = s.status; *bytes = B * 4; break;
    case AVSIM_LATCH: *p = s.latch; *bytes = B * 4; break;
    case AVSIM_QACC: *p = s.qacc; *bytes = B * d.nv * 4; break;
    case AVSIM_XPOS: *p = s.xpos; *bytes = B * 3 * d.nbody * 4; break;
    case AVSIM_QFRC_BIAS: *p = s.qfrc_bias; *bytes = B * d.nv * 4; break;
    case AVSIM_QACC_SMOOTH: *p = s.qacc_smooth; *bytes = B * d.nv * 4; break;
    case AVSIM_MASS_DIAG: *p = s.mass_diag; *bytes = B * d.nv * 4; break;
    case AVSIM_ENV_CYCLES: *p = s.env_cycles; *bytes = B * 8; break;
    case AVSIM_FC_KEY: *p = s.fc_key; *bytes = B * (AV_NCON + AV_NSC) * 4; break;
    case AVSIM_FC_N: *p = s.fc_n; *bytes = B * 2 * 4; break;
    case AVSIM_FC_VAL: *p = s.fc_val; *bytes = B * (AV_NCON * 6 + AV_NSC) * 4; break;
    case AVSIM_SOLVER_STAT: *p = s.nw_stat; *bytes = B * 4 * 4; break;
    default: return fail(AVSIM_ERR_ARG, "unknown field");
    }
    return AVSIM_OK;
}

__global__ void avsim_success_kernel(const int *reward, int max_reward, int *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = reward[i] == max_reward;
}

extern "C" int avsim_get(avsim_batch *b, int field, void *dst_dev) {
    if (!b || !dst_dev) return fail(AVSIM_ERR_ARG, "avsim_get: null argument");
    CU(cudaSetDevice(b->model->device));
    if (field == AVSIM_SUCCESS) {
        int n = b->st.num_envs;
        avsim_success_kernel<<<(n + 255) / 256, 256, 0, b->stream>>>(b->st.reward, b->model->dm.max_reward, (int *)dst_dev, n);
        b->launches++;
        CU(cudaGetLastError());
        return AVSIM_OK;
    }
    void *p;
    size_t bytes;
    int rc = field_ptr(b, field, &p, &bytes);
    if (rc) return rc;
    CU(cudaMemcpyAsync(dst_dev, p, bytes, cudaMemcpyDeviceToDevice, b->stream));
    return AVSIM_OK;
}

extern "C" int avsim_set(avsim_batch *b, int field, const void *src_dev) {
    if (!b || !src_dev) return fail(AVSIM_ERR_ARG, "avsim_set: null argument");
    if (field != AVSIM_QPOS && field != AVSIM_QVEL && field != AVSIM_CTRL && field != AVSIM_WARMSTART && field != AVSIM_LATCH &&
        field != AVSIM_FC_KEY && field != AVSIM_FC_N && field != AVSIM_FC_VAL)
        return fail(AVSIM_ERR_ARG, "avsim_set: field is read-only");
    CU(cudaSetDevice(b->model->device));
    void *p;
    size_t bytes;
    int rc = field_ptr(b, field, &p, &bytes);
    if (rc) return rc;
    CU(cudaMemcpyAsync(p, src_dev, bytes, cudaMemcpyDeviceToDevice, b->stream));
    return AVSIM_OK;
}

extern "C" int avsim_step_host(avsim_batch *b, const float *action_host, int nsubsteps, float *agent_pos_host, int32_t *reward_host, int32_t *status_host) {
    if (!b || !action_host) return fail(AVSIM_ERR_ARG, "avsim_step_host: null argument");
    CU(cudaSetDevice(b->model->device));
    const DevModel &d = b->model->dm;
    size_t na = (size_t)b->st.num_envs * d.nj_obs;
    memcpy(b->h_action, action_host, na * sizeof(float));
    CU(cudaMemcpyAsync(b->d_action, b->h_action, na * sizeof(float), cudaMemcpyHostToDevice, b->stream));
    int rc = avsim_step(b, b->d_action, nsubsteps);
    if (rc) return rc;
    CU(cudaMemcpyAsync(b->h_agent, b->st.agent_pos, na * sizeof(float), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaMemcpyAsync(b->h_reward, b->st.reward, b->st.num_envs * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    if (status_host) CU(cudaMemcpyAsync(b->h_status, b->st.status, b->st.num_envs * sizeof(int32_t), cudaMemcpyDeviceToHost, b->stream));
    CU(cudaStreamSynchronize(b->stream));
    if (agent_pos_host) memcpy(agent_pos_host, b->h_agent, na * sizeof(float));
    if (reward_host) memcpy(reward_host, b->h_reward, b->st.num_envs * sizeof(int32_t));
    if (status_host) memcpy(status_host, b->h_status, b->st.num_envs * sizeof(int32_t));
    return AVSIM_OK;
}

// render: replaces physics.render(h, w, camera_id) per configured camera (reference env.py:180-188, 195-200)
static int render_impl(avsim_batch *b, const int *cam_ids_host, int ncam, int H, int W, uint8_t *dst_dev, int id_mode);
extern "C" int avsim_render(avsim_batch *b, const int *cam_ids_host, int ncam, int H, int W, uint8_t *dst_dev) {
    return render_impl(b, cam_ids_host, ncam, H, W, dst_dev, 0);
}
// test hook: same rays, but every pixel carries the index of the geom it sees in all three channels (255 = background)
extern "C" int avsim_render_ids(avsim_batch *b, const int *cam_ids_host, int ncam, int H, int W, uint8_t *dst_dev) {
    return render_impl(b, cam_ids_host, ncam, H, W, dst_dev, 1);
}
static int render_impl(avsim_batch *b, const int *cam_ids_host, int ncam, int H, int W, uint8_t *dst_dev, int id_mode) {
    if (!b || !cam_ids_host || !dst_dev || ncam < 1 || ncam > 8 || H < 1 || W < 4 || (W % 4))
        return fail(AVSIM_ERR_ARG, "avsim_render: bad arguments (1..8 cameras, width a multiple of 4)");
    const DevModel &d = b->model->dm;
    for (int k = 0; k < ncam; k++)
        if (cam_ids_host[k] < 0 || cam_ids_host[k] >= d.ncam) return fail(AVSIM_ERR_ARG, "avsim_render: camera id out of range");
    CU(cudaSetDevice(b->model->device));
    size_t n = (size_t)b->st.num_envs;
    if (!b->d_rpose) {
        if (!dalloc(b, &b->d_rpose, n * (d.ngeom + d.ncam) * 12) || !dalloc(b, &b->d_camids, 8) ||
            !dalloc(b, &b->d_rrect, n * 8 * d.ngeom * 3))
            return fail(AVSIM_ERR_CUDA, "avsim_render: device allocation failed");
    }
    CU(cudaMemcpyAsync(b->d_camids, cam_ids_host, ncam * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    avsim_render_prep_kernel<<<b->fwd_grid, 32, sizeof(EnvS), b->stream>>>(d, b->st, b->d_rpose, b->d_rrect, d.ncam, d.cam_body, d.cam_pos,
                                                                            d.cam_quat, d.cam_fovy, b->d_camids, ncam, H, W);
    CU(cudaGetLastError());
    dim3 grid(((W + AV_RT_W - 1) / AV_RT_W) * ((H + AV_RT_H - 1) / AV_RT_H), ncam, (unsigned)n);
    avsim_render_kernel<<<grid, AV_RT_THREADS, 0, b->stream>>>(d, b->d_rpose, b->d_rrect, d.geom_rgba, d.geom_visible, d.cam_fovy, b->d_camids, ncam,
                                                               d.ncam, H, W, dst_dev, id_mode);
    b->launches += 2;
    CU(cudaGetLastError());
    return AVSIM_OK;
}

// device-resident preprocess_observation (reference lerobot/lerobot/common/envs/utils.py:37-50): u8 [n][H][W][3] -> f32 [n][3][H][W] / 255
extern "C" int avsim_pixels_to_float(const uint8_t *src_dev, int64_t n_images, int H, int W, float *dst_dev, int device, void *stream) {
    if (n_images < 0 || H < 1 || W < 1) return fail(AVSIM_ERR_ARG, "avsim_pixels_to_float: bad arguments");
    if (n_images == 0) return AVSIM_OK;                       // an empty batch has no buffers to look at
    if (!src_dev || !dst_dev) return fail(AVSIM_ERR_ARG, "avsim_pixels_to_float: null buffer");
    CU(cudaSetDevice(device));
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const long long hw = (long long)H * W, cap = (long long)sms * 8;        // 8 resident 256-thread blocks per SM
    if (hw % 4 == 0 && ((uintptr_t)src_dev % 4) == 0 && ((uintptr_t)dst_dev % 16) == 0) {
        long long nquads = n_images * (hw / 4), per = (long long)AV_OBS_THREADS * AV_OBS_UNROLL;
        long long grid = (nquads + per - 1) / per;
        avsim_pixels_to_float_kernel<<<(unsigned)(grid < cap ? grid : cap), AV_OBS_THREADS, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint32_t *>(src_dev), dst_dev, nquads, (int)(hw / 4));
    } else {
        long long nvals = n_images * 3 * hw, grid = (nvals + AV_OBS_THREADS - 1) / AV_OBS_THREADS;
        avsim_pixels_to_float_any_kernel<<<(unsigned)(grid < cap ? grid : cap), AV_OBS_THREADS, 0, (cudaStream_t)stream>>>(src_dev, dst_dev, nvals, (int)hw);
    }
    CU(cudaGetLastError());
    return AVSIM_OK;
}

extern "C" int avsim_stage_cycles(uint64_t *out_host, int n, int reset) {
#ifdef AVSIM_PROFILE
    unsigned long long h[PF_N] = {0};
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(h, g_prof, sizeof h));
    for (int k = 0; k < n && k < PF_N; k++) out_host[k] = h[k];
    if (reset) { unsigned long long z[PF_N] = {0}; CU(cudaMemcpyToSymbol(g_prof, z, sizeof z)); }
    return PF_N;
#else
    (void)out_host; (void)n; (void)reset;
    return fail(AVSIM_ERR_ARG, "avsim_stage_cycles: library built without -DAVSIM_PROFILE");
#endif
}

extern "C" int64_t avsim_launch_count(const avsim_batch *b) { return b ? b->launches : 0; }
extern "C" int avsim_launch_shape(const avsim_batch *b, int out[6]) {
    if (!b || !out) return fail(AVSIM_ERR_ARG, "avsim_launch_shape: null argument");
    out[0] = (b->split && b->st.solver == AVSIM_SOLVER_NEWTON) ? 1 : 0;
    out[1] = out[0] ? b->ngroups : 1; out[2] = b->warps; out[3] = b->envw; out[4] = b->solve_warps; out[5] = b->sms;
    return AVSIM_OK;
}

// ------------------------------------------------------------------ IK entry points
extern "C" int avsim_diffik(const avsim_model *m, int arm, const float *q, const float *pos, const float *quat, int n,
                            const avsim_diffik_params *p, float *q_out, void *stream) {
    if (!m || !q || !pos || !quat || !p || !q_out || arm < 0 || arm > 2 || n < 0) return fail(AVSIM_ERR_ARG, "avsim_diffik: bad arguments");
    CU(cudaSetDevice(m->device));
    if (n == 0) return AVSIM_OK;
    DiffIKParams dp;
    dp.k_pos = p->k_pos; dp.k_ori = p->k_ori; dp.damping = p->damping; dp.max_angvel = p->max_angvel;
    dp.dt = p->integration_dt; dp.iterations = p->iterations;
    for (int k = 0; k < 7; k++) { dp.k_null[k] = p->k_null[k]; dp.q0[k] = p->q0[k]; }
    avsim_diffik_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(m->dm, arm, q, pos, quat, n, dp, q_out);
    CU(cudaGetLastError());
    return AVSIM_OK;
}
extern "C" int avsim_gradik(const avsim_model *m, int arm, const float *q, const float *pos, const float *quat, int n,
                            const avsim_gradik_params *p, float *q_out, void *stream) {
    if (!m || !q || !pos || !quat || !p || !q_out || arm < 0 || arm > 2 || n < 0) return fail(AVSIM_ERR_ARG, "avsim_gradik: bad arguments");
    CU(cudaSetDevice(m->device));
    if (n == 0) return AVSIM_OK;
    GradIKParams gp;
    gp.step_size = p->step_size; gp.min_cost_delta = p->min_cost_delta; gp.position_weight = p->position_weight;
    gp.rotation_weight = p->rotation_weight; gp.position_threshold = p->position_threshold;
    gp.rotation_threshold = p->rotation_threshold; gp.max_pos_diff = p->max_pos_diff; gp.max_rot_diff = p->max_rot_diff;
    gp.joint_p = p->joint_p; gp.max_iterations = p->max_iterations;
    for (int k = 0; k < 7; k++) { gp.center_w[k] = p->joint_center_weight[k]; gp.disp_w[k] = p->joint_displacement_weight[k]; }
    // up to AV_GRADIK_GROUP_MAX problems: 16 lanes per problem (latency: 3 FK chains per iteration instead of 16); above, one
    // thread per problem (throughput: every lane busy).  AVSIM_GRADIK_GROUP=0/1 forces one form (tests compare the two).
    const char *fe = getenv("AVSIM_GRADIK_GROUP");
    const int force = fe ? atoi(fe) : -1;
    if (force == 1 || (force < 0 && n <= AV_GRADIK_GROUP_MAX))
        avsim_gradik_group_kernel<<<(n + 7) / 8, 128, 0, (cudaStream_t)stream>>>(m->dm, arm, q, pos, quat, n, gp, q_out);
    else
        avsim_gradik_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(m->dm, arm, q, pos, quat, n, gp, q_out);
    CU(cudaGetLastError());
    return AVSIM_OK;
}
extern "C" int avsim_transform(int op, const double *a_dev, const double *b_dev, int n, double p0, double p1, double *out_dev, int device, void *stream) {
    if (op < 0 || op >= XF_NOPS || n < 0) return fail(AVSIM_ERR_ARG, "avsim_transform: unknown operator or negative count");
    if (n == 0) return AVSIM_OK;
    bool two = op == XF_ANGULAR_ERROR || op == XF_LIMIT_POSE || op == XF_WITHIN_POSE;
    if (!a_dev || !out_dev || (two && !b_dev)) return fail(AVSIM_ERR_ARG, "avsim_transform: null buffer");
    CU(cudaSetDevice(device));
    avsim_transform_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(op, a_dev, b_dev, n, p0, p1, out_dev);
    CU(cudaGetLastError());
    return AVSIM_OK;
}
extern "C" int avsim_fk(const avsim_model *m, int arm, const float *q, int n, float *T_out, void *stream) {
    if (!m || !q || !T_out || arm < 0 || arm > 2 || n < 0) return fail(AVSIM_ERR_ARG, "avsim_fk: bad arguments");
    CU(cudaSetDevice(m->device));
    if (n == 0) return AVSIM_OK;
    avsim_fk_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(m->dm, arm, q, n, T_out);
    CU(cudaGetLastError());
    return AVSIM_OK;
}
extern "C" int avsim_jac(const avsim_model *m, int arm, const float *q, int n, float *J_out, void *stream) {
    if (!m || !q || !J_out || arm < 0 || arm > 2 || n < 0) return fail(AVSIM_ERR_ARG, "avsim_jac: bad arguments");
    CU(cudaSetDevice(m->device));
    if (n == 0) return AVSIM_OK;
    avsim_jac_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(m->dm, arm, q, n, J_out);
    CU(cudaGetLastError());
    return AVSIM_OK;
}
