// avsim_kernels.cuh -- __global__ entry points: env step (nsub substeps + trailing position pass + reward),
// forward (stage dumps for parity tests) and reset.  One warp (= one 32-thread block) per environment;
// the grid is a persistent multiple of the SM count and loops over environments.
#pragma once
#include "avsim_step.cuh"

extern __shared__ float4 av_smem_raw[];
#ifndef AV_SHARED
#define AV_SHARED __shared__   // the host emulation harness maps it to `static` (one variable per block)
#endif

// once per kernel and warp: the two mbarriers of the solver's contact-block double buffer
__device__ inline void env_pipe_init(EnvS &S, int lane) {
#if AV_BULK_PREFETCH
    if (lane == 0) {
        mbar_init(&S.mbar[0], 1);
        mbar_init(&S.mbar[1], 1);
        S.cuse[0] = S.cuse[1] = 0;
    }
    __syncwarp();
#else
    (void)S; (void)lane;
#endif
}

// per-launch constants of a slice: world-welded bodies keep their compile-time pose; the per-tree dof table
__device__ inline void env_consts(const DevModel &m, EnvS &S, int lane) {
    for (int b = lane; b < m.nbody; b += 32)
        if (m.body_tree[b] < 0) {
            st3(S.xpos + 3 * b, ld3(m.body_xpos0 + 3 * b));
            Q4 q = ldq(m.body_xquat0 + 4 * b);
            stq(S.xquat + 4 * b, q);
        }
    if (lane < m.ntree) S.tree_pk[lane] = m.tree_dofadr[lane] | (m.tree_dofnum[lane] << 6);
}
__device__ inline void env_load(const DevModel &m, const BatchState &B, EnvS &S, int env, int lane) {
    for (int i = lane; i < m.nq; i += 32) S.qpos[i] = B.qpos[(size_t)env * m.nq + i];
    for (int i = lane; i < m.nv; i += 32) { S.qvel[i] = B.qvel[(size_t)env * m.nv + i]; S.warm[i] = B.warm[(size_t)env * m.nv + i]; }
    for (int i = lane; i < m.nu; i += 32) S.ctrl[i] = B.ctrl[(size_t)env * m.nu + i];
    env_consts(m, S, lane);
    // padding of the per-tree 8x8 blocks must be finite: block_apply multiplies it by exact zeros
    for (int i = lane; i < AV_MBLK; i += 32) S.Minv[i] = 0.f;
    if (lane == 0) { S.status = 0; S.ncon = 0; S.nsc = 0; }
    __syncwarp();
}
// split pipeline: 0 = the noslip sweeps run at the start of the next substep kernel launch (lockstep blocks: their cost is the same
// for every environment), 1 = at the end of the solve kernel on the warp that holds the solution.  Measured: 40.1 vs 45.7 ms
// per env.step at B = 4096 (profiles/r2_sweeps.txt) -- in the solve kernel the sweeps stretch the trip of every phase-locked block.
#ifndef AV_NOSLIP_IN_SOLVE
#define AV_NOSLIP_IN_SOLVE 0
#endif
// ---- the head record of a slice <-> its image in global memory (split pipeline).  On the device one lane issues ONE bulk
// asynchronous copy (TMA engine: cp.async.bulk, UBLKCP in SASS) per direction and the warp waits on an mbarrier / bulk group;
// -DAV_TMA_HEAD=0 falls back to 128-bit loads and stores by all lanes (the A/B is in profiles/).
#ifndef AV_TMA_HEAD
#define AV_TMA_HEAD 1
#endif
__device__ inline void head_load(EnvS &S, const float *img, int n_floats, unsigned long long *bar, unsigned &phase, int lane) {
#if defined(__CUDA_ARCH__) && AV_TMA_HEAD
    __syncwarp();
    if (lane == 0) {
        fence_proxy_async();   // this warp's earlier generic accesses to the slice are ordered before the async write
        bulk_g2s(&S, img, (unsigned)n_floats * 4u, bar);
    }
    while (!mbar_wait(bar, phase & 1u)) {}
    phase++;
#else
    (void)bar; (void)phase;
    const float4 *src = reinterpret_cast<const float4 *>(img);
    float4 *dst = reinterpret_cast<float4 *>(&S);
    for (int i = lane; i < n_floats / 4; i += 32) dst[i] = src[i];
    __syncwarp();
#endif
}
__device__ inline void head_store(const EnvS &S, float *img, int n_floats, int lane) {
#if defined(__CUDA_ARCH__) && AV_TMA_HEAD
    __syncwarp();
    if (lane == 0) {
        fence_proxy_async();   // the slice was written through the generic proxy
        bulk_s2g(img, &S, (unsigned)n_floats * 4u);
        bulk_commit_wait();    // the slice may be overwritten (next task) once the engine has read it
    }
    __syncwarp();
#else
    const float4 *src = reinterpret_cast<const float4 *>(&S);
    float4 *dst = reinterpret_cast<float4 *>(img);
    for (int i = lane; i < n_floats / 4; i += 32) dst[i] = src[i];
    __syncwarp();
#endif
}
__device__ inline void env_store(const DevModel &m, const BatchState &B, EnvS &S, int env, int lane) {
    for (int i = lane; i < m.nq; i += 32) B.qpos[(size_t)env * m.nq + i] = S.qpos[i];
    for (int i = lane; i < m.nv; i += 32) { B.qvel[(size_t)env * m.nv + i] = S.qvel[i]; B.warm[(size_t)env * m.nv + i] = S.warm[i]; }
    for (int i = lane; i < m.nu; i += 32) B.ctrl[(size_t)env * m.nu + i] = S.ctrl[i];
}

// with_reward: publish the staged reward (and SewNeedle's latch) of the contacts just found -- env.step and physics.forward()
// + get_reward() (reference scripts/check_dataset_reward.py:35-38 read the reward right after set_qpos); the forward pass that
// reset launches keeps reward 0 and the cleared latch instead.  after_solve: the contact forces in S.c_f are valid (dumped).
__device__ inline void env_outputs(const DevModel &m, const BatchState &B, EnvS &S, const float *scratch, int env, int lane, bool with_reward,
                                   bool after_solve) {
    int latch = B.latch[env];
    int r = stage_reward(m, S, lane, latch);
    if (!with_reward) { r = B.reward[env]; latch = B.latch[env]; }
    for (int k = lane; k < m.nj_obs; k += 32) {
        float q = S.qpos[m.obs_qadr[k]];
        if (k == 6 || k == 13) q = (q - m.act_ctrl_lo[k]) / (m.act_ctrl_hi[k] - m.act_ctrl_lo[k]);
        B.agent_pos[(size_t)env * m.nj_obs + k] = q;
    }
    // numerical blow-up guard: flag and leave the state for the host-side auto-reset
    int bad = 0;
    for (int i = lane; i < m.nq; i += 32) bad |= !(fabsf(S.qpos[i]) < 1e6f);
    bad = __any_sync(AV_FULL, bad);
    if (lane == 0) {
        B.reward[env] = r;
        B.latch[env] = latch;
        B.ncon[env] = S.ncon;
        B.status[env] = S.status | (bad ? 1 : 0);
    }
    for (int c = lane; c < S.ncon; c += 32) {
        float *o = B.contacts + ((size_t)env * AV_NCON + c) * 16;
        int info = S.c_info[c];
        const float *geo = scratch + c * AV_CBLK + AV_CB_GEO;
        o[0] = geo[12];
        o[1] = geo[0]; o[2] = geo[1]; o[3] = geo[2];
        o[4] = geo[3]; o[5] = geo[4]; o[6] = geo[5];
        o[7] = (float)(info & 0xff); o[8] = (float)((info >> 8) & 0xff); o[9] = (float)((info >> 16) & 0xf);
        o[10] = (float)((info >> 20) & 1); o[11] = after_solve ? S.c_f[6 * c] : 0.f;
        o[12] = o[13] = o[14] = o[15] = 0.f;
    }
}

__device__ __forceinline__ FCache env_fcache(const BatchState &B, int env) {
    FCache fc;
    fc.key = B.fc_key + (size_t)env * (AV_NCON + AV_NSC);
    fc.val = B.fc_val + (size_t)env * (AV_NCON * 6 + AV_NSC);
    fc.n = B.fc_n + 2 * (size_t)env;
    fc.mode = B.solver == 1 ? 3 : B.warm_mode;   // 3: primal solver, no dual warm start (rows skip it; Lc = noslip factor)
    return fc;
}

// per-environment solver statistics of one launch: Newton iterations summed over the substeps, the largest scaled gradient
// a solve ended with, the most iterations one solve took, solves that hit the iteration cap (B.nw_stat, 4 floats per env)
struct NwStat {
    float iters = 0.f, grad = 0.f, worst = 0.f, capped = 0.f;
    __device__ __forceinline__ void add(int ni, float gr, int cap) {
        iters += (float)ni; grad = fmaxf(grad, gr); worst = fmaxf(worst, (float)ni); capped += ni >= cap ? 1.f : 0.f;
    }
    __device__ __forceinline__ void store(const BatchState &B, int env, int lane) const {
        if (lane == 0) {
            float *o = B.nw_stat + 4 * (size_t)env;
            o[0] = iters; o[1] = grad; o[2] = worst; o[3] = capped;
        }
    }
};

__device__ __forceinline__ void env_forward(const DevModel &m, const BatchState &B, EnvS &S, float *scratch, int lane, Prof &pf, const FCache &fc, NwStat &nw, float *bias_out) {
    stage_kinematics(m, S, lane);
    pf.mark(PF_KIN, lane);
    stage_inertia(m, S, lane);
    pf.mark(PF_INERTIA, lane);
    stage_collision(m, S, scratch, lane, B.multiccd != 0, pf);
    stage_smooth(m, S, lane);
    pf.mark(PF_SMOOTH, lane);
    if (bias_out)   // parity dump (forward kernel): the Newton solve reuses S.qfrc_bias
        for (int i = lane; i < m.nv; i += 32) bias_out[i] = S.qfrc_bias[i];
    stage_rows_scalar(m, S, lane, fc);
    pf.mark(PF_ROWS_S, lane);
    stage_rows_contact(m, S, scratch, lane, fc);
    pf.mark(PF_ROWS_C, lane);
    if (B.solver == 1) {   // Newton on the primal, then the noslip sweeps on its forces
        float gr = 0.f;
        int ni = stage_newton(m, S, scratch, lane, B.newton_iters, B.newton_ls, B.newton_tol, gr, pf);
        nw.add(ni, gr, B.newton_iters);
        stage_solve(m, S, scratch, lane, 0, B.noslip_iters, 3, true);
    } else
        stage_solve(m, S, scratch, lane, B.solver_iters, B.noslip_iters, B.warm_mode);   // physics.forward(): the force cache is read, not updated
    pf.mark(PF_SOLVE, lane);
}

// env.step: ctrl write (reference env.py:204-215), nsub x mj_step (env.py:218), trailing position pass, reward.
//
// Block = W warps = W environments, each with its own EnvS slice, moving through the pipeline STAGE BY STAGE in lockstep
// (__syncthreads between stages).  Why: the kernel is bound by instruction fetch, not by data -- ncu showed
// 'no_instruction' as the dominant stall and throughput independent of the number of resident warps, because every
// free-running warp streams its own ~200 KB of code per substep through the SM's instruction cache.  In lockstep all
// warps of the SM execute the same stage, so a fetched line serves W warps.
// Work distribution: a queue (B.order / B.queue) instead of a static stride, handed out W entries at a time.
// Environments differ ~5x in cost; avsim_order_kernel sorts the queue by the cycles each environment took in the
// previous step (costliest first), which (a) fills the tail of the launch with cheap environments and (b) puts
// environments of similar cost into the same block, so the lockstep barriers wait for little.
// Block shape (measured, profiles/r1_summary.md): blockDim.y = 16 warps, of which the first B.env_warps = 11 own an
// environment slice; the other 5 are HELPER warps without shared-memory state that only pull pooled narrowphase items
// (16 warps x 128 registers fill the SM's register file).  The slice count is chosen by the L1 it leaves, not by the
// shared memory it fills: shared memory and L1 split one 256 KB array in steps (... 132, 164, 196, 228 KB); 11 slices of
// 11.8 KB stay inside the 132 KB step (~124 KB of L1 for hull vertices and contact blocks), 12-13 slices need the 164 KB
// step and run 6 % slower although more environments are resident, 14 need 196 KB and lose another 4 %.
// Collision of a lockstep block: phase A per environment, then the MPR runs of ALL the block's environments are pooled and
// every warp (also the ones without an environment) pulls items -- first the unperturbed runs, then the perturbed
// (multiccd) runs of the pairs that touch -- then phase C per environment.
__device__ __forceinline__ void pool_round(const DevModel &m, EnvS &S, int lane, int warp, int W, int my_items, int per_pair,
                                           int *s_next, int *s_off, float *const *s_scr, unsigned long long *s_cost) {
    if (lane == 0) s_off[warp + 1] = my_items;
    if (warp == 0 && lane == 0) { *s_next = 0; s_off[0] = 0; }
    __syncthreads();
    if (warp == 0 && lane == 0)
        for (int w = 0; w < W; w++) s_off[w + 1] += s_off[w];
    __syncthreads();
    const int total = s_off[W];
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(s_next, 1);
        item = __shfl_sync(AV_FULL, item, 0);
        if (item >= total) break;
        int w = 0;
        while (item >= s_off[w + 1]) w++;
        const EnvS &Se = *(reinterpret_cast<const EnvS *>(av_smem_raw) + w);
        int local = item - s_off[w];
        long long ti = clock64();
        if (per_pair == 1) collide_item(m, Se, s_scr[w], local, 0, lane);
        else collide_item(m, Se, s_scr[w], Se.cand_c[local >> 2], 1 + (local & 3), lane);
        if (lane == 0) atomicAdd(&s_cost[w], (unsigned long long)(clock64() - ti));
    }
    __syncthreads();
}

__device__ __forceinline__ void block_collision(const DevModel &m, const BatchState &B, EnvS &S, float *scratch, int lane, int warp, int W,
                                                bool active, Prof &pf, long long &own) {
    AV_SHARED int s_next, s_off[AV_MAX_WARPS + 1];
    AV_SHARED float *s_scr[AV_MAX_WARPS];
    AV_SHARED unsigned long long s_cost[AV_MAX_WARPS];   // cycles spent (by any warp) on each environment's pooled items
    const bool multiccd = B.multiccd != 0;
    long long t0 = clock64();
    if (active) stage_collision_a(m, S, scratch, lane, pf);
    if (active) own += clock64() - t0;
    if (lane == 0) { s_scr[warp] = scratch; s_cost[warp] = 0ull; }
    pool_round(m, S, lane, warp, W, active ? S.nkeep : 0, 1, &s_next, s_off, s_scr, s_cost);
    if (active) collide_select(m, S, scratch, lane, multiccd);
    pool_round(m, S, lane, warp, W, active ? 4 * S.ncand_c : 0, 4, &s_next, s_off, s_scr, s_cost);
    t0 = clock64();
    if (active) stage_collision_c(m, S, scratch, lane, multiccd, pf);
    if (active) own += (clock64() - t0) + (B.key_pooled ? (long long)s_cost[warp] : 0ll);
    if (B.sync >= 2) __syncthreads();
}

// `own` accumulates the cycles this warp spends on ITS environment, excluding the waits at the lockstep barriers: the
// queue is sorted by it.  (Sorting by wall cycles per environment degenerates within a few dozen steps: in lockstep all
// environments of a block finish together, so every one of them inherits the cost of the slowest and the order stops
// reflecting the environments themselves -- measured as a drift from 47 to 58 ms/step.)
#define AV_STAGE_SYNC(call)                       \
    do {                                          \
        if (active) {                             \
            long long t0_ = clock64();            \
            call;                                 \
            own += clock64() - t0_;               \
        }                                         \
        if (B.sync >= 2) __syncthreads();         \
    } while (0)

__global__ void __launch_bounds__(32 * AV_MAX_WARPS, 1) avsim_step_kernel(const __grid_constant__ DevModel m, const __grid_constant__ BatchState B,
                                                                          const float *__restrict__ action, int nsub) {
    const int lane = threadIdx.x, warp = threadIdx.y, W = blockDim.y, EW = B.env_warps;
    EnvS &S = *(reinterpret_cast<EnvS *>(av_smem_raw) + (warp < EW ? warp : 0));   // helper warps (warp >= EW) never touch S
    AV_SHARED int s_base;
    Prof pf;
    if (warp < EW) env_pipe_init(S, lane);
    for (;;) {
        if (warp == 0 && lane == 0) s_base = atomicAdd(B.queue, 1);
        __syncthreads();
        const int task = s_base;
        __syncthreads();
        // task -> slice of the sorted order: the first heavy_tasks tasks take heavy_warps environments each, the rest W
        const int nenv = task < B.heavy_tasks ? B.heavy_warps : EW;
        const int base = task < B.heavy_tasks ? task * B.heavy_warps : B.heavy_tasks * B.heavy_warps + (task - B.heavy_tasks) * EW;
        if (base >= B.num_envs) break;
        const bool active = warp < nenv && base + warp < B.num_envs;
        const int env = active ? B.order[base + warp] : 0;
        float *scratch = B.scratch + (size_t)env * AV_SCRATCH_FLOATS;
        const FCache fc = env_fcache(B, env);
        pf.start();
        long long own = 0;
        NwStat nw;
        if (active) {
            env_load(m, B, S, env, lane);
            pf.mark(PF_LOAD, lane);
            if (action && lane < m.nj_obs) {
                float a = action[(size_t)env * m.nj_obs + lane];
                if (lane == 6 || lane == 13) a = a * (m.act_ctrl_hi[lane] - m.act_ctrl_lo[lane]) + m.act_ctrl_lo[lane];
                S.ctrl[lane] = a;
            }
            __syncwarp();
        }
        for (int s = 0; s < nsub; s++) {
            if (B.sync == 1) __syncthreads();
            AV_STAGE_SYNC(stage_kinematics(m, S, lane); pf.mark(PF_KIN, lane); stage_inertia(m, S, lane); pf.mark(PF_INERTIA, lane));
            block_collision(m, B, S, scratch, lane, warp, W, active, pf, own);
            AV_STAGE_SYNC(stage_smooth(m, S, lane); pf.mark(PF_SMOOTH, lane); stage_rows_scalar(m, S, lane, fc); pf.mark(PF_ROWS_S, lane);
                          stage_rows_contact(m, S, scratch, lane, fc); pf.mark(PF_ROWS_C, lane));
            const int sweeps = B.solver == 1 ? 0 : B.solver_iters;   // Newton replaces the regularised sweeps; noslip follows either
            if (B.solver == 1) AV_STAGE_SYNC(float gr = 0.f; int ni = stage_newton(m, S, scratch, lane, B.newton_iters, B.newton_ls, B.newton_tol, gr, pf);
                                             nw.add(ni, gr, B.newton_iters));
            AV_STAGE_SYNC(stage_solve_begin(m, S, scratch, lane, B.solver == 1 ? 3 : B.warm_mode));
            for (int it = 0; it < sweeps + B.noslip_iters; it++) {   // sync 3: the sweeps in lockstep too
                if (active) {
                    long long t0 = clock64();
                    solve_sweep(m, S, scratch, lane, it >= sweeps, B.solver == 1);
                    own += clock64() - t0;
                }
                if (B.sync >= 3) __syncthreads();
            }
            if (active && B.solver != 1) stage_cache_store(m, S, lane, fc);
            pf.mark(PF_SOLVE, lane);
            if (B.sync == 2) __syncthreads();
            AV_STAGE_SYNC(stage_integrate(m, S, lane); pf.mark(PF_INTEGRATE, lane));
        }
        AV_STAGE_SYNC(stage_kinematics(m, S, lane); pf.mark(PF_KIN, lane));
        block_collision(m, B, S, scratch, lane, warp, W, active, pf, own);
        if (active) {
            env_store(m, B, S, env, lane);
            env_outputs(m, B, S, scratch, env, lane, true, false);
            pf.mark(PF_OUT, lane);
            if (lane == 0) B.env_cycles[env] = own;
            nw.store(B, env, lane);
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ split pipeline: substep kernel + solver kernel
// One env.step = for s in 0..nsub: avsim_substep_kernel(s); if (s < nsub) avsim_solve_kernel();   (2 nsub + 1 launches)
//
// Why two kernels.  The Newton solve takes 1..17 iterations depending on how many contacts change zone in that substep.  Inside
// the lockstep blocks of the fused kernel every solve costs its block the iterations of the block's SLOWEST environment
// (measured: 11 ms per step per mean Newton iteration; profiles/r2_summary.md), and letting the warps free-run through the
// whole pipeline instead thrashes the instruction cache (AVSIM_SYNC=0: 73 vs 46 ms).  The solver kernel holds only the solver's
// code and its own block shape (phase-locked warps, see avsim_solve_kernel): one warp = one environment pulled from a
// cost-sorted queue, an environment that needs 17 iterations holds one warp for 17 trips and delays nobody else.  The substep
// kernel keeps the lockstep blocks + pooled narrowphase for everything else (whose cost per environment is far more uniform).
// What crosses the kernel boundary is the slice's head record (AV_HEAD_FLOATS floats, 6.7 KB: state, mass matrix and its
// block inverse, smooth forces, scalar rows, contact ids) -- one bulk copy per direction through an L2-resident image -- and
// the contact blocks, which already live in the global scratch.  Per env.step that is 2 x 20 x 6.7 KB x 2 = 0.5 MB per
// environment of L2 traffic (2.2 GB per launch pair at B = 4096, ~0.4 ms at L2 bandwidth).
__global__ void __launch_bounds__(32 * AV_MAX_WARPS, 1) avsim_substep_kernel(const __grid_constant__ DevModel m, const __grid_constant__ BatchState B,
                                                                             const float *__restrict__ action, int s, int nsub) {
    const int lane = threadIdx.x, warp = threadIdx.y, W = blockDim.y, EW = B.env_warps;
    EnvS &S = *(reinterpret_cast<EnvS *>(av_smem_raw) + (warp < EW ? warp : 0));   // helper warps (warp >= EW) never touch S
    AV_SHARED int s_base;
    AV_SHARED unsigned long long s_hbar[AV_MAX_WARPS];
    unsigned hphase = 0;
    Prof pf;
    if (lane == 0) mbar_init(&s_hbar[warp], 1);
    if (blockIdx.x == 0 && warp == 0 && lane == 0) *B.queue_b = 0;   // rewind the solver kernel's queue (it is not running now)
    __syncthreads();
    for (;;) {
        if (warp == 0 && lane == 0) s_base = atomicAdd(B.queue, 1);
        __syncthreads();
        const int task = s_base;
        __syncthreads();
        const int nenv = task < B.heavy_tasks ? B.heavy_warps : EW;
        const int base = task < B.heavy_tasks ? task * B.heavy_warps : B.heavy_tasks * B.heavy_warps + (task - B.heavy_tasks) * EW;
        if (base >= B.num_envs) break;
        const bool active = warp < nenv && base + warp < B.num_envs;
        const int env = active ? B.order[base + warp] : 0;
        float *scratch = B.scratch + (size_t)env * AV_SCRATCH_FLOATS;
        float *img = B.heads + (size_t)env * AV_HEADX_FLOATS;
        const FCache fc = env_fcache(B, env);
        pf.start();
        long long own = 0;
        if (active) {
            if (s == 0) {
                env_load(m, B, S, env, lane);
                if (action && lane < m.nj_obs) {
                    float a = action[(size_t)env * m.nj_obs + lane];
                    if (lane == 6 || lane == 13) a = a * (m.act_ctrl_hi[lane] - m.act_ctrl_lo[lane]) + m.act_ctrl_lo[lane];
                    S.ctrl[lane] = a;
                }
                if (lane == 0) { float *o = B.nw_stat + 4 * (size_t)env; o[0] = o[1] = o[2] = o[3] = 0.f; B.env_cycles_b[env] = 0; }
            } else {
                head_load(S, img, AV_HEADX_FLOATS, &s_hbar[warp], hphase, lane);   // state + the solver's acc and forces
                env_consts(m, S, lane);
            }
            pf.mark(PF_LOAD, lane);
            __syncwarp();
        }
        if (s > 0) {   // [the noslip sweeps on the solver's forces,] then the integrator
#if !AV_NOSLIP_IN_SOLVE
            AV_STAGE_SYNC(stage_solve_begin(m, S, scratch, lane, 3));
            for (int it = 0; it < B.noslip_iters; it++) AV_STAGE_SYNC(solve_sweep(m, S, scratch, lane, true, true));
            pf.mark(PF_SOLVE, lane);
#endif
            AV_STAGE_SYNC(stage_integrate(m, S, lane); pf.mark(PF_INTEGRATE, lane));
        }
        AV_STAGE_SYNC(stage_kinematics(m, S, lane); pf.mark(PF_KIN, lane); if (s < nsub) { stage_inertia(m, S, lane); pf.mark(PF_INERTIA, lane); });
        block_collision(m, B, S, scratch, lane, warp, W, active, pf, own);
        if (s < nsub) {
            AV_STAGE_SYNC(stage_smooth(m, S, lane); pf.mark(PF_SMOOTH, lane); stage_rows_scalar(m, S, lane, fc); pf.mark(PF_ROWS_S, lane);
                          stage_rows_contact(m, S, scratch, lane, fc); pf.mark(PF_ROWS_C, lane));
            if (active) {
                head_store(S, img, AV_HEAD_FLOATS, lane);
                if (lane == 0) B.env_cycles[env] = (s == 0 ? 0 : B.env_cycles[env]) + own;
            }
        } else if (active) {   // trailing position pass done: publish the step
            env_store(m, B, S, env, lane);
            env_outputs(m, B, S, scratch, env, lane, true, false);
            pf.mark(PF_OUT, lane);
            if (lane == 0) B.env_cycles[env] += own;
            __syncwarp();
        }
    }
}

// The Newton solve of one substep for every environment.  One block per SM, its warps PHASE-LOCKED: every trip of the loop runs
// the phases fetch / gradient / Hessian / direction / line search once, with a block barrier after each, and a warp takes part in
// a phase when its environment needs it.  A warp whose environment has converged publishes the forces, stores the record and
// fetches the next environment from the queue at the top of the next trip -- so an environment that needs 17 iterations holds
// one warp for 17 trips and nobody else.  Why phase-locked: free-running warps (the first version of this kernel) spread over
// 100 KB of solver code and ncu showed 'no_instruction' as 6.4 of the 12.4 stall cycles per issued instruction, at any
// occupancy (profiles/r2_summary.md row 2; the final kernel: profiles/r2_step_kernels_ncu.txt) -- instruction fetch, not arithmetic, bounded it; in lockstep one fetched line
// serves all the warps of the SM.  The noslip sweeps that follow the solve cost the same for every environment, so they run in
// the substep kernel's lockstep blocks (start of the next launch), not here.
#ifndef AV_NW_SYNC
#define AV_NW_SYNC 4   // phase barriers per Newton trip besides the one at the top (sweep: profiles/r2_sweeps.txt)
#endif
__global__ void __launch_bounds__(32 * AV_MAX_WARPS, 1) avsim_solve_kernel(const __grid_constant__ DevModel m, const __grid_constant__ BatchState B) {
    const int lane = threadIdx.x, warp = threadIdx.y;
    EnvS &S = *reinterpret_cast<EnvS *>(reinterpret_cast<char *>(av_smem_raw) + (size_t)warp * AV_SOLVER_SLICE_BYTES);
    AV_SHARED unsigned long long s_hbar[AV_MAX_WARPS];
    unsigned hphase = 0;
    if (lane == 0) mbar_init(&s_hbar[warp], 1);
    if (blockIdx.x == 0 && warp == 0 && lane == 0) *B.queue = 0;   // rewind the substep kernel's queue
    __syncthreads();
    Newton nw;
    bool run = false, drained = false;
    int env = 0;
    float *scratch = nullptr, *img = nullptr;
    long long t0 = 0;
    Prof pf;   // -DAVSIM_PROFILE: own cycles per phase + cycles spent waiting at the phase barriers
    pf.start();
    for (;;) {
        if (!run && !drained) {   // ---- fetch: next environment of the cost-sorted queue, record in, start point
            int idx = 0;
            if (lane == 0) idx = atomicAdd(B.queue_b, 1);
            idx = __shfl_sync(AV_FULL, idx, 0);
            if (idx < B.num_envs) {
                env = B.order_b[idx];
                scratch = B.scratch + (size_t)env * AV_SCRATCH_FLOATS;
                img = B.heads + (size_t)env * AV_HEADX_FLOATS;
                t0 = clock64();
                head_load(S, img, AV_HEAD_FLOATS, &s_hbar[warp], hphase, lane);
                nw.init(m, S, scratch, lane);
                run = true;
            } else
                drained = true;
        }
        pf.mark(PF_NW_INIT, lane);
        if (!__syncthreads_or(run ? 1 : 0)) break;
        pf.mark(PF_NW_WAIT, lane);
        bool stop = false;
        if (run) stop = nw.grad(m, S, scratch, lane, B.newton_iters, B.newton_tol);
        pf.mark(PF_NW_GRAD, lane);
#if AV_NW_SYNC >= 4
        __syncthreads();
        pf.mark(PF_NW_WAIT, lane);
#endif
        if (run && !stop) nw.hess(m, S, scratch, lane);
        pf.mark(PF_NW_HESS, lane);
#if AV_NW_SYNC >= 2
        __syncthreads();
        pf.mark(PF_NW_WAIT, lane);
#endif
        if (run && !stop) stop = nw.dir(m, S, lane, B.newton_tol);
        pf.mark(PF_NW_CHOL, lane);
#if AV_NW_SYNC >= 3
        __syncthreads();
        pf.mark(PF_NW_WAIT, lane);
#endif
        if (run && !stop) stop = nw.search(m, S, scratch, lane, B.newton_ls);
        if (run && stop) {   // converged (or out of descent in fp32, or at the cap): forces out, record back to the image
            float gr = 0.f;
            nw.publish(S, lane, gr);
#if AV_NOSLIP_IN_SOLVE
            // experiment (see the macro): the noslip sweeps on the warp that still holds the solution in shared memory
            stage_solve_begin(m, S, scratch, lane, 3);
            for (int it = 0; it < B.noslip_iters; it++) solve_sweep(m, S, scratch, lane, true, true);
#endif
            if (lane == 0) {
                float *o = B.nw_stat + 4 * (size_t)env;
                o[0] += (float)nw.it; o[1] = fmaxf(o[1], gr); o[2] = fmaxf(o[2], (float)nw.it); o[3] += nw.it >= B.newton_iters ? 1.f : 0.f;
            }
            head_store(S, img, AV_HEADX_FLOATS, lane);   // head (acc, scalar-row forces, status) + contact forces / multipliers
            if (lane == 0) B.env_cycles_b[env] += clock64() - t0;
            run = false;
        }
        pf.mark(PF_NW_LS, lane);
        __syncthreads();
        pf.mark(PF_NW_WAIT, lane);
    }
}

// physics.forward(): all stages, no integration; dumps stage outputs for the parity tests.  mask (nullable) selects envs.
__global__ void __launch_bounds__(32, AV_MIN_BLOCKS) avsim_forward_kernel(const __grid_constant__ DevModel m, const __grid_constant__ BatchState B,
                                                                         const uint8_t *__restrict__ mask, int publish_reward) {
    int lane = threadIdx.x;
    EnvS &S = *reinterpret_cast<EnvS *>(av_smem_raw);
    env_pipe_init(S, lane);
    for (int env = blockIdx.x; env < B.num_envs; env += gridDim.x) {
        if (mask && !mask[env]) continue;
        Prof pf;
        pf.start();
        env_load(m, B, S, env, lane);
        float *scratch = B.scratch + (size_t)env * AV_SCRATCH_FLOATS;
        NwStat nw;
        env_forward(m, B, S, scratch, lane, pf, env_fcache(B, env), nw, B.qfrc_bias + (size_t)env * m.nv);
        nw.store(B, env, lane);
        for (int i = lane; i < m.nv; i += 32) {
            B.qacc[(size_t)env * m.nv + i] = S.qacc_smooth[i] + S.acc[i];
            B.qacc_smooth[(size_t)env * m.nv + i] = S.qacc_smooth[i];
            int t = m.dof_tree[i], dl = i - m.tree_dofadr[t];
            B.mass_diag[(size_t)env * m.nv + i] = S.M[t * AV_MTRI + av_mtri(dl, dl)];
        }
        for (int i = lane; i < 3 * m.nbody; i += 32) B.xpos[(size_t)env * 3 * m.nbody + i] = S.xpos[i];
        env_outputs(m, B, S, scratch, env, lane, publish_reward != 0, true);
        __syncwarp();
    }
}

// Sorts environments by the cycles they took in the previous step (descending) into B.order and rewinds the queue.
// Bitonic sort of (cycles, env) keys in shared memory.  Up to `chunk` (<= 8192, a power of two) environments one block does
// it all (n2 = num_envs rounded up to a power of two).  Larger batches: block c sorts environments [c chunk, (c+1) chunk) and
// the sorted chunks are interleaved -- order = rank 0 of every chunk, rank 1 of every chunk, ... -- which is heaviest-first
// to within one chunk's spread; that is all the queue needs (blocks of similar cost, the long environments early).
__global__ void avsim_order_kernel(BatchState B, int n2, int chunk) {
    extern __shared__ unsigned long long av_keys[];
    const int c = blockIdx.x, nch = gridDim.x, e0 = c * chunk;
    const int n = min(chunk, B.num_envs - e0);
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        unsigned long long cy = i < n ? (unsigned long long)B.env_cycles[e0 + i] : 0ull;
        if (cy > 0xffffffffffull) cy = 0xffffffffffull;
        // key = cycles << 20 | (0xfffff - env): descending sort keeps ties in ascending env order; padding sorts last
        av_keys[i] = i < n ? ((cy << 20) | (unsigned long long)(0xfffff - i)) : 0ull;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                int l = i ^ j;
                if (l > i) {
                    unsigned long long a = av_keys[i], b = av_keys[l];
                    bool desc = (i & k) == 0;
                    if (desc ? a < b : a > b) { av_keys[i] = b; av_keys[l] = a; }
                }
            }
            __syncthreads();
        }
    // only the last chunk can be short (L environments): ranks below L interleave over all chunks, the rest over nch - 1
    const int L = B.num_envs - (nch - 1) * chunk;
    for (int r = threadIdx.x; r < n; r += blockDim.x) {
        int pos = r < L ? r * nch + c : L * nch + (r - L) * (nch - 1) + c;
        B.order[pos] = e0 + 0xfffff - (int)(av_keys[r] & 0xfffff);
    }
    if (c == 0 && threadIdx.x == 0) *B.queue = 0;
}
__global__ void avsim_identity_order_kernel(BatchState B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B.num_envs) B.order[i] = i;
    if (i == 0) *B.queue = 0;
}

// ---- Philox4x32-10 counter RNG for device-side reset draws
__device__ inline void philox4x32(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// reset: home pose, open fingers, ctrl = home, object placement (reference env.py:228-249 + task resets); thread = env
__global__ void avsim_reset_kernel(DevModel m, BatchState B, const uint8_t *__restrict__ mask, const float *__restrict__ free_pos,
                                   const float *__restrict__ home /*[21]*/) {
    int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= B.num_envs) return;
    if (mask && !mask[env]) return;
    float *qpos = B.qpos + (size_t)env * m.nq, *qvel = B.qvel + (size_t)env * m.nv, *ctrl = B.ctrl + (size_t)env * m.nu,
          *warm = B.warm + (size_t)env * m.nv;
    for (int i = 0; i < m.nq; i++) qpos[i] = m.qpos0[i];
    for (int i = 0; i < m.nv; i++) { qvel[i] = 0.f; warm[i] = 0.f; }
    for (int k = 0; k < 21; k++) { qpos[m.obs_qadr[k]] = home[k]; ctrl[k] = home[k]; }
    float open_l = m.act_ctrl_hi[6], open_r = m.act_ctrl_hi[13];
    qpos[m.finger_qadr[0]] = qpos[m.finger_qadr[1]] = open_l;
    qpos[m.finger_qadr[2]] = qpos[m.finger_qadr[3]] = open_r;
    ctrl[6] = open_l; ctrl[13] = open_r;
    int episode = B.episode[env];
    for (int k = 0; k < m.nfree; k++) {
        int qa = m.free_qadr[k];
        float p[3];
        if (free_pos) {
            for (int a = 0; a < 3; a++) p[a] = free_pos[((size_t)env * m.nfree + k) * 3 + a];
        } else {
            int src = m.reset_draw[k];   // TubeTransfer's tube1 re-uses the ball's draw (env.py:713-716)
            uint32_t c[4] = {(uint32_t)env, (uint32_t)episode, (uint32_t)src, 0x41564c4fu};
            philox4x32(c, (uint32_t)B.seed, (uint32_t)(B.seed >> 32));
            for (int a = 0; a < 3; a++) {
                float u = (c[a] >> 8) * (1.0f / 16777216.0f);
                p[a] = m.reset_lo[3 * k + a] + u * (m.reset_hi[3 * k + a] - m.reset_lo[3 * k + a]);
            }
        }
        qpos[qa] = p[0]; qpos[qa + 1] = p[1]; qpos[qa + 2] = p[2];
        qpos[qa + 3] = 1.f; qpos[qa + 4] = qpos[qa + 5] = qpos[qa + 6] = 0.f;
    }
    B.fc_n[2 * env] = B.fc_n[2 * env + 1] = 0;   // a fresh episode has no force history
    B.latch[env] = 0;
    B.reward[env] = 0;
    B.status[env] = 0;
    B.episode[env] = episode + 1;
}
