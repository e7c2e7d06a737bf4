// avsim_emu.cpp -- TEST/DEBUG HARNESS ONLY (see cuda_emu.h): the step / forward / reset kernels of
// av_aloha_b200/csrc compiled for the host and driven lane-by-lane through 32 cooperative fibers, exposed through
// a tiny C interface that tests/test_emu_parity.py loads with ctypes.  It exists so the kernel *source* can be
// checked against the oracle in the GPU-less build container; it is never loaded by the package.
#include "cuda_emu.h"

namespace emu {
Warp W;
dim3_emu g_blockIdx = {0, 0, 0}, g_gridDim = {1, 1, 1}, g_blockDim = {32, 1, 1};
static void trampoline() {
    W.body();
    W.done[W.cur] = true;
    swapcontext(&W.ctx[W.cur], &W.main);
}
void run_block(int block, int grid, const std::function<void()> &body) {
    static bool init = false;
    if (!init) {
        for (int i = 0; i < 32; i++) W.stack[i] = (char *)malloc(1 << 20);
        init = true;
    }
    g_blockIdx.x = block; g_gridDim.x = grid;
    W.body = body; W.arrived = 0; W.gen = 0;
    for (int i = 0; i < 32; i++) {
        W.done[i] = false;
        getcontext(&W.ctx[i]);
        W.ctx[i].uc_stack.ss_sp = W.stack[i];
        W.ctx[i].uc_stack.ss_size = 1 << 20;
        W.ctx[i].uc_link = &W.main;
        makecontext(&W.ctx[i], trampoline, 0);
    }
    for (;;) {
        bool all = true;
        for (int i = 0; i < 32; i++)
            if (!W.done[i]) {
                all = false;
                W.cur = i;
                swapcontext(&W.main, &W.ctx[i]);
            }
        if (all) break;
    }
}
}  // namespace emu

#include "../../av_aloha_b200/csrc/avsim_kernels.cuh"
#include "../../av_aloha_b200/csrc/avsim_ik.cuh"
#include "../../av_aloha_b200/csrc/avsim_model_pack.h"

unsigned long long av_keys[8192];
float4 av_smem_raw[(sizeof(EnvS) + 15) / 16 + 1];
#include <cstddef>

struct EmuBatch {
    avpack::PackedModel pk;
    BatchState st;
    std::vector<float> f[16];
    std::vector<int> iv[8];
    std::vector<long long> cyc;
    std::vector<int> order, queue, fckey, fcn, order_b, queue_b;
    std::vector<float> fcval, heads;
    std::vector<long long> cyc_b;
    int split = 1;   // Newton: 1 = substep + solve kernels (large batches on the device), 0 = the fused step kernel (small batches)
};

extern "C" {
EmuBatch *emu_create(const char *path, int num_envs) {
    EmuBatch *b = new EmuBatch();
    if (!avpack::pack_model(path, b->pk)) { fprintf(stderr, "emu_create: %s\n", b->pk.error.c_str()); delete b; return nullptr; }
    b->pk.relocate(b->pk.P.fdata.data(), b->pk.P.idata.data(), b->pk.hull4.data());
    const DevModel &d = b->pk.dm;
    BatchState &s = b->st;
    memset(&s, 0, sizeof s);
    size_t B = num_envs;
    s.num_envs = num_envs; s.seed = 1234; s.solver_iters = 50; s.noslip_iters = d.noslip_iterations; s.multiccd = d.multiccd;
    auto F = [&](int k, size_t n) { b->f[k].assign(n, 0.f); return b->f[k].data(); };
    auto I = [&](int k, size_t n) { b->iv[k].assign(n, 0); return b->iv[k].data(); };
    s.qpos = F(0, B * d.nq); s.qvel = F(1, B * d.nv); s.ctrl = F(2, B * d.nu); s.warm = F(3, B * d.nv);
    s.agent_pos = F(4, B * d.nj_obs); s.contacts = F(5, B * AV_NCON * 16); s.qacc = F(6, B * d.nv);
    s.xpos = F(7, B * 3 * d.nbody); s.qfrc_bias = F(8, B * d.nv); s.qacc_smooth = F(9, B * d.nv);
    s.mass_diag = F(10, B * d.nv); s.scratch = F(11, B * AV_SCRATCH_FLOATS); s.nw_stat = F(12, B * 4);
    b->cyc.assign(B, 0); s.env_cycles = b->cyc.data();
    b->fckey.assign(B * (AV_NCON + AV_NSC), 0); b->fcn.assign(B * 2, 0); b->fcval.assign(B * (AV_NCON * 6 + AV_NSC), 0.f);
    s.fc_key = b->fckey.data(); s.fc_n = b->fcn.data(); s.fc_val = b->fcval.data(); s.warm_mode = 1;
    s.solver = 1; s.newton_iters = 30; s.newton_ls = 20; s.newton_tol = 3e-7f;   // the library's defaults (avsim_create)
    b->order.resize(B); b->queue.assign(1, 0); s.order = b->order.data(); s.queue = b->queue.data();
    b->heads.assign(B * AV_HEADX_FLOATS, 0.f); s.heads = b->heads.data();
    b->order_b.resize(B); for (size_t i = 0; i < B; i++) b->order_b[i] = (int)i;
    b->queue_b.assign(1, 0); b->cyc_b.assign(B, 0);
    s.order_b = b->order_b.data(); s.queue_b = b->queue_b.data(); s.env_cycles_b = b->cyc_b.data();
    s.env_warps = 1; s.key_pooled = 1;   // the emulated block is one warp = one environment
    s.reward = I(0, B); s.status = I(1, B); s.latch = I(2, B); s.ncon = I(3, B); s.episode = I(4, B);
    return b;
}
void emu_destroy(EmuBatch *b) { delete b; }
void emu_set_warmstart(EmuBatch *b, int mode) { b->st.warm_mode = mode; }
void emu_set_split(EmuBatch *b, int split) { b->split = split; }
void emu_set_solver(EmuBatch *b, int solver, int max_iter, int ls_iter, float tol) {
    b->st.solver = solver; b->st.newton_iters = max_iter; b->st.newton_ls = ls_iter; b->st.newton_tol = tol;
}
void emu_set_options(EmuBatch *b, int iters, int noslip, int multiccd) {
    b->st.solver_iters = iters;
    b->st.noslip_iters = noslip >= 0 ? noslip : b->pk.dm.noslip_iterations;
    b->st.multiccd = multiccd >= 0 ? multiccd : b->pk.dm.multiccd;
}
int emu_dim(EmuBatch *b, int what) {
    const DevModel &d = b->pk.dm;
    switch (what) { case 0: return d.nq; case 1: return d.nv; case 2: return d.nu; case 3: return d.nbody; case 4: return d.nj_obs; case 5: return d.nfree; }
    return -1;
}
float *emu_f(EmuBatch *b, int k) {
    BatchState &s = b->st;
    float *p[] = {s.qpos, s.qvel, s.ctrl, s.warm, s.agent_pos, s.contacts, s.qacc, s.xpos, s.qfrc_bias, s.qacc_smooth, s.mass_diag, s.nw_stat};
    return p[k];
}
int *emu_i(EmuBatch *b, int k) {
    BatchState &s = b->st;
    int *p[] = {s.reward, s.status, s.latch, s.ncon, s.episode};
    return p[k];
}
static void emu_forward_impl(EmuBatch *b, int publish_reward) {
    for (int e = 0; e < b->st.num_envs; e++)
        emu::run_block(e, b->st.num_envs, [&]() { avsim_forward_kernel(b->pk.dm, b->st, nullptr, publish_reward); });
}
void emu_forward(EmuBatch *b) { emu_forward_impl(b, 1); }
void emu_step(EmuBatch *b, const float *action, int nsub) {
    int n2 = 1;
    while (n2 < b->st.num_envs) n2 <<= 1;
    emu::run_block(0, 1, [&]() { avsim_order_kernel(b->st, n2, 8192); });   // same queue order as the device
    if (b->st.solver == 1 && b->split) {   // split pipeline, as avsim_step launches it above ~2 500 environments
        for (int s = 0; s <= nsub; s++) {
            emu::run_block(0, 1, [&]() { avsim_substep_kernel(b->pk.dm, b->st, action, s, nsub); });
            if (s < nsub) emu::run_block(0, 1, [&]() { avsim_solve_kernel(b->pk.dm, b->st); });
        }
    } else
        emu::run_block(0, 1, [&]() { avsim_step_kernel(b->pk.dm, b->st, action, nsub); });   // one block drains the queue
}
// the queue-order kernel on its own: n environments with the given cycle counts, sorted in chunks of `chunk` (one block each)
void emu_order(EmuBatch *b, const long long *cycles, int n, int chunk, int *order_out) {
    BatchState st = b->st;
    int queue = -1;
    st.num_envs = n;
    st.env_cycles = const_cast<long long *>(cycles);
    st.order = order_out;
    st.queue = &queue;
    int n2 = 1, nch = (n + chunk - 1) / chunk;
    while (n2 < (n < chunk ? n : chunk)) n2 <<= 1;
    for (int c = 0; c < nch; c++) emu::run_block(c, nch, [&]() { avsim_order_kernel(st, n2, chunk); });
}
void emu_reset_masked(EmuBatch *b, const uint8_t *mask, const float *free_pos) {
    int nb = (b->st.num_envs + 31) / 32;
    for (int blk = 0; blk < nb; blk++)
        emu::run_block(blk, nb, [&]() { avsim_reset_kernel(b->pk.dm, b->st, mask, free_pos, AV_HOME); });
    for (int e = 0; e < b->st.num_envs; e++)
        if (!mask || mask[e]) emu::run_block(e, b->st.num_envs, [&]() { avsim_forward_kernel(b->pk.dm, b->st, mask, 0); });
}
void emu_reset(EmuBatch *b, const float *free_pos) { emu_reset_masked(b, nullptr, free_pos); }
// test hook: stage_reward (the CUDA source) on an explicit contact list of geom id pairs; returns the reward, writes the latch back
int emu_reward_from_pairs(EmuBatch *b, const int *pairs, int n, int latch_in, int *latch_out) {
    if (n > AV_NCON) return -1;
    int reward = -1, latch = latch_in;
    emu::run_block(0, 1, [&]() {
        EnvS &S = *reinterpret_cast<EnvS *>(av_smem_raw);
        int lane = threadIdx.x;
        if (lane == 0) {
            S.ncon = n;
            for (int c = 0; c < n; c++) S.c_info[c] = pairs[2 * c] | (pairs[2 * c + 1] << 8) | (3 << 16);
        }
        __syncwarp();
        int l = latch_in;
        int r = stage_reward(b->pk.dm, S, lane, l);
        if (lane == 0) { reward = r; latch = l; }
    });
    if (latch_out) *latch_out = latch;
    return reward;
}
// IK kernels (thread per problem): arm 0 left, 1 right, 2 middle
void emu_fk(EmuBatch *b, int arm, const float *q, int n, float *T_out) {
    int nb = (n + 31) / 32;
    for (int blk = 0; blk < nb; blk++) emu::run_block(blk, nb, [&]() { avsim_fk_kernel(b->pk.dm, arm, q, n, T_out); });
}
void emu_jac(EmuBatch *b, int arm, const float *q, int n, float *J_out) {
    int nb = (n + 31) / 32;
    for (int blk = 0; blk < nb; blk++) emu::run_block(blk, nb, [&]() { avsim_jac_kernel(b->pk.dm, arm, q, n, J_out); });
}
void emu_diffik(EmuBatch *b, int arm, const float *q, const float *pos, const float *quat, int n, const DiffIKParams *p, float *out) {
    int nb = (n + 31) / 32;
    for (int blk = 0; blk < nb; blk++) emu::run_block(blk, nb, [&]() { avsim_diffik_kernel(b->pk.dm, arm, q, pos, quat, n, *p, out); });
}
void emu_transform(int op, const double *a, const double *b, int n, double p0, double p1, double *out) {
    int nb = (n + 31) / 32;
    for (int blk = 0; blk < nb; blk++) emu::run_block(blk, nb, [&]() { avsim_transform_kernel(op, a, b, n, p0, p1, out); });
}
void emu_gradik(EmuBatch *b, int arm, const float *q, const float *pos, const float *quat, int n, const GradIKParams *p, float *out) {
    int nb = (n + 31) / 32;
    for (int blk = 0; blk < nb; blk++) emu::run_block(blk, nb, [&]() { avsim_gradik_kernel(b->pk.dm, arm, q, pos, quat, n, *p, out); });
}
}
