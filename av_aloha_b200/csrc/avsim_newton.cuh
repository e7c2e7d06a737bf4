// avsim_newton.cuh -- K6 (second form): Newton on the primal problem, one warp per environment, fp32.
//
// The reference never sets `solver=` (reference gym_guided_vision/gym_guided_vision/assets/aloha_sim.xml:4-6), so
// `physics.step` (reference env.py:218) solves its constraints with MuJoCo's default Newton solver, tolerance 1e-8, and then
// runs 3 noslip sweeps [third-party engine; the algorithm is restated from its published description, SURVEY.md Appendix A,
// and from oracle/avsim_oracle.c solve_newton, which is checked against the dual PGS path run to convergence].
//
// Unconstrained convex problem in acc = qacc - qacc_smooth:
//     minimise  0.5 acc' M acc  +  sum over constraint blocks  s(J acc + b),      b = J qacc_smooth - aref
// s = convex conjugate of the block's dual problem: equality rows 0.5 D y^2; friction loss Huber; limits one-sided quadratic;
// elliptic contacts in three zones -- top (inside the polar cone: 0), bottom (inside the cone: sum 0.5 D_k y_k^2), middle
// (0.5 Dm (N - mu T)^2 with N = mu y_0, T = |friction_k y_k|, mu = friction_0 sqrt(R_1/R_0), Dm = D_0 / (mu^2 (1 + mu^2))).
// Every iteration: gradient g = M acc - J' f, Hessian H = M + J' Hc J (dense nv x nv, packed lower triangle in shared memory),
// in-warp Cholesky, search direction p = -H^-1 g, exact line search on the piecewise-smooth 1-D restriction (safeguarded
// Newton on its derivative; per block a handful of polynomial coefficients in registers, lane = constraint block).
//
// Why this shape on the GPU: the dual has 6 rows per contact (nefc ~ 110) and block Gauss-Seidel needs hundreds of sweeps with
// impratio = 100; the primal has nv = 35..41 unknowns, a warm-started Newton step is a 35 x 35 Cholesky plus ~20 rank-6 updates,
// and 2-3 of them reach the converged answer.  Everything that is summed into shared memory is summed by ONE lane per address
// per instruction (no float atomics), so results are bit-reproducible from run to run.
#pragma once

// (i << 4 | j) of the e-th entry of a row-major lower triangle; the first k (k + 1) / 2 entries are the triangle of order k
__constant__ unsigned char av_tri_ij[136] = {
    0, 16, 17, 32, 33, 34, 48, 49, 50, 51, 64, 65, 66, 67, 68, 80, 81, 82, 83, 84, 85, 96, 97, 98, 99, 100, 101, 102, 112, 113, 114,
    115, 116, 117, 118, 119, 128, 129, 130, 131, 132, 133, 134, 135, 136, 144, 145, 146, 147, 148, 149, 150, 151, 152, 153, 160, 161,
    162, 163, 164, 165, 166, 167, 168, 169, 170, 176, 177, 178, 179, 180, 181, 182, 183, 184, 185, 186, 187, 192, 193, 194, 195, 196,
    197, 198, 199, 200, 201, 202, 203, 204, 208, 209, 210, 211, 212, 213, 214, 215, 216, 217, 218, 219, 220, 221, 224, 225, 226, 227,
    228, 229, 230, 231, 232, 233, 234, 235, 236, 237, 238, 240, 241, 242, 243, 244, 245, 246, 247, 248, 249, 250, 251, 252, 253, 254,
    255};

__device__ __forceinline__ int nw_tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// parameters of one elliptic contact: regularisers of rows [0 | 1,2 | 3 | 4,5], friction [f0 f0 f1 f2 f2], regularised cone slope
struct ConeP {
    float R[6], sc[6], mu, Dm;
};
__device__ __forceinline__ void cone_load(const float *blk, ConeP &p) {
    float Rn = ldblk1(blk + AV_CB_PAR), Rf = ldblk1(blk + AV_CB_PAR + 1), Rt = ldblk1(blk + AV_CB_PAR + 2), Rr = ldblk1(blk + AV_CB_PAR + 3);
    float f0 = ldblk1(blk + AV_CB_PAR + 4), f1 = ldblk1(blk + AV_CB_PAR + 5), f2 = ldblk1(blk + AV_CB_PAR + 6);
    p.R[0] = Rn; p.R[1] = p.R[2] = Rf; p.R[3] = Rt; p.R[4] = p.R[5] = Rr;
    p.mu = f0 * sqrtf(Rf / Rn);
    p.sc[0] = p.mu; p.sc[1] = p.sc[2] = f0; p.sc[3] = f1; p.sc[4] = p.sc[5] = f2;
    p.Dm = 1.0f / (Rn * p.mu * p.mu * (1.0f + p.mu * p.mu));
}
// zone of a contact at the constraint-space value y: 0 top (no force), 1 bottom (quadratic), 2 middle (cone surface).
// Dead rows of a condim-3 contact carry J = 0 and b = 0, so their y is 0 and they drop out of every sum.
__device__ __forceinline__ int cone_zone(const float (&y)[6], const ConeP &p, float (&U)[6], float &T, float &e) {
    float T2 = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) U[k] = p.sc[k] * y[k];
#pragma unroll
    for (int k = 1; k < 6; k++) T2 += U[k] * U[k];
    T = sqrtf(T2);
    e = U[0] - p.mu * T;
    if (e >= 0.f) return 0;
    if (p.mu * U[0] + T <= 0.f) return 1;
    return 2;
}
// force = -ds/dy and cost s(y)
__device__ __forceinline__ float cone_force(const float (&y)[6], const ConeP &p, float (&f)[6]) {
    float U[6], T, e;
    int z = cone_zone(y, p, U, T, e);
    float cost = 0.f;
    if (z == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) f[k] = 0.f;
    } else if (z == 1) {
#pragma unroll
        for (int k = 0; k < 6; k++) { f[k] = -__fdividef(y[k], p.R[k]); cost -= 0.5f * f[k] * y[k]; }
    } else {
        float a = p.Dm * e * p.mu, iT = __fdividef(1.0f, fmaxf(T, 1e-30f));
        f[0] = -a;
#pragma unroll
        for (int k = 1; k < 6; k++) f[k] = a * U[k] * iT * p.sc[k];
        cost = 0.5f * p.Dm * e * e;
    }
    return cost;
}
// block Hessian d2s/dy2, packed lower triangle (21); returns the zone
__device__ __forceinline__ int cone_hess(const float (&y)[6], const ConeP &p, float (&Hc)[21]) {
    float U[6], T, e;
    int z = cone_zone(y, p, U, T, e);
#pragma unroll
    for (int k = 0; k < 21; k++) Hc[k] = 0.f;
    if (z == 1) {
#pragma unroll
        for (int k = 0; k < 6; k++) Hc[TRI(k, k)] = __fdividef(1.0f, p.R[k]);
    } else if (z == 2) {
        float iT = __fdividef(1.0f, fmaxf(T, 1e-30f)), u[6];
#pragma unroll
        for (int k = 1; k < 6; k++) u[k] = U[k] * iT;
        float w = -e * p.mu * iT;   // > 0 in the middle zone
        Hc[0] = p.Dm * p.sc[0] * p.sc[0];
#pragma unroll
        for (int k = 1; k < 6; k++) {
            Hc[TRI(k, 0)] = -p.Dm * p.mu * u[k] * p.sc[k] * p.sc[0];
#pragma unroll
            for (int l = 1; l <= k; l++)
                Hc[TRI(k, l)] = p.Dm * (p.mu * p.mu * u[k] * u[l] + w * ((k == l ? 1.0f : 0.0f) - u[k] * u[l])) * p.sc[k] * p.sc[l];
        }
    }
    return z;
}

// The solver's running constraint-space value jar = J acc + b of contact c (6 floats) lives in the contact's scratch block, in the
// slots of the two frame tangents: the row assembly is the last reader of the tangents, the outputs read only position / normal / distance.
#define AV_CB_JAR (AV_CB_GEO + 6)
__device__ __forceinline__ float *nw_jar(float *scratch, int c) { return scratch + c * AV_CBLK + AV_CB_JAR; }
__device__ __forceinline__ const float *nw_jar(const float *scratch, int c) { return scratch + c * AV_CBLK + AV_CB_JAR; }

// J of contact c (this lane's three rows at its column) times a joint-space vector in shared memory: six row sums, every lane
__device__ __forceinline__ void nw_rows(const float *blk, int tr, const float *x, int lane, float (&res)[6]) {
    int col = lane & 15, half = lane >> 4, dof = tr_dof(tr, col);
    const float *J = blk + AV_CB_J + (3 * half) * AV_JW + col;
    block_rows(ldblk1(J), ldblk1(J + AV_JW), ldblk1(J + 2 * AV_JW), dof >= 0 ? x[dof] : 0.f, half, res);
}
__device__ __forceinline__ float sel6(const float (&v)[6], int k) {
    float r = v[0];
#pragma unroll
    for (int i = 1; i < 6; i++) r = (k == i) ? v[i] : r;
    return r;
}
// (M x)_i for this lane's dof i (x in shared memory)
__device__ __forceinline__ float nw_mrow(const DevModel &m, const EnvS &S, const float *x, int i) {
    int t = m.dof_tree[i], d0 = m.tree_dofadr[t], nt = m.tree_dofnum[t];
    const float *Mb = S.M + t * AV_MTRI;
    float s = 0.f;
    for (int k = 0; k < nt; k++) s += Mb[av_mtri(i - d0, k)] * x[d0 + k];
    return s;
}
// scalar row r at acc: constraint-space value y, force (clamped), whether the clamp is inactive
__device__ __forceinline__ float sc_jx(const EnvS &S, int r, const float *x) {
    return S.sc_c1[r] * x[S.sc_dof1[r]] + (S.sc_dof2[r] >= 0 ? S.sc_c2[r] * x[S.sc_dof2[r]] : 0.f);
}
__device__ __forceinline__ float sc_force(const EnvS &S, int r, float y, bool &free_) {
    float f = -__fdividef(y, S.sc_R[r]);
    free_ = f > S.sc_lo[r] && f < S.sc_hi[r];
    return fminf(fmaxf(f, S.sc_lo[r]), S.sc_hi[r]);
}

// jar = J x + b of every contact
__device__ __forceinline__ void nw_jar_all(EnvS &S, float *scratch, const float *x, int lane) {
    for (int c = 0; c < S.ncon; c++) {
        if ((S.c_info[c] >> 20) & 1) continue;
        float *blk = scratch + c * AV_CBLK;
        float res[6];
        nw_rows(blk, c_tr(S, c), x, lane, res);
        if (lane < 6) blk[AV_CB_JAR + lane] = sel6(res, lane) + ldblk1(blk + AV_CB_B + lane);
    }
    __syncwarp();
}
// constraint cost at the stored jar (b_only: at acc = 0, where jar = b) + scalar rows at x
__device__ __forceinline__ float nw_cost(const EnvS &S, const float *scratch, const float *x, bool b_only, int lane) {
    float cost = 0.f;
    for (int c = lane; c < S.ncon; c += 32) {
        if ((S.c_info[c] >> 20) & 1) continue;
        const float *blk = scratch + c * AV_CBLK;
        ConeP p;
        cone_load(blk, p);
        float y[6], f[6];
#pragma unroll
        for (int k = 0; k < 6; k++) y[k] = ldblk1(blk + (b_only ? AV_CB_B : AV_CB_JAR) + k);
        cost += cone_force(y, p, f);
    }
    if (lane < S.nsc) {
        bool fr;
        float y = (b_only ? 0.f : sc_jx(S, lane, x)) + S.sc_b[lane], f = sc_force(S, lane, y, fr);
        cost -= f * (y + 0.5f * S.sc_R[lane] * f);
    }
    return warp_sum(cost);
}

// per-lane line-search state of one contact: N(a) = N0 + a N1, T^2(a) = T0 + a T1 + a^2 T2, bottom-zone quadratic Q1, Q2
struct LsCon {
    float N0, N1, T0, T1, T2, Q1, Q2, mu, Dm, v[6];
    bool live;
};
__device__ __forceinline__ void ls_con_init(LsCon &L, const EnvS &S, const float *scratch, int c, const float (&v)[6]) {
    L.live = c < S.ncon && !((S.c_info[c] >> 20) & 1);
    L.N0 = L.N1 = L.T0 = L.T1 = L.T2 = L.Q1 = L.Q2 = 0.f; L.mu = 1.f; L.Dm = 0.f;
#pragma unroll
    for (int k = 0; k < 6; k++) L.v[k] = v[k];
    if (!L.live) return;
    ConeP p;
    cone_load(scratch + c * AV_CBLK, p);
    L.mu = p.mu; L.Dm = p.Dm;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        float y = ldblk1(scratch + c * AV_CBLK + AV_CB_JAR + k), a = p.sc[k] * y, b = p.sc[k] * v[k], D = __fdividef(1.0f, p.R[k]);
        if (k == 0) { L.N0 = a; L.N1 = b; }
        else { L.T0 += a * a; L.T1 += 2.f * a * b; L.T2 += b * b; }
        L.Q1 += D * y * v[k]; L.Q2 += D * v[k] * v[k];
    }
}
__device__ __forceinline__ void ls_con_eval(const LsCon &L, float a, float &d1, float &d2) {
    if (!L.live) return;
    float N = L.N0 + a * L.N1, T = sqrtf(fmaxf(L.T0 + a * (L.T1 + a * L.T2), 0.f)), e = N - L.mu * T;
    if (e >= 0.f) return;
    if (L.mu * N + T <= 0.f) { d1 += L.Q1 + a * L.Q2; d2 += L.Q2; return; }
    float iT = __fdividef(1.0f, fmaxf(T, 1e-30f)), Tp = 0.5f * (L.T1 + 2.f * a * L.T2) * iT, Tpp = (L.T2 - Tp * Tp) * iT;
    float ep = L.N1 - L.mu * Tp;
    d1 += L.Dm * e * ep;
    d2 += L.Dm * (ep * ep - e * L.mu * Tpp);
}

// in-place Cholesky of the packed lower triangle (order n); the diagonal keeps 1 / L_jj.  Left-looking, column by column: the
// rows at and below the pivot are spread over the lanes -- one lane per row while more than 16 rows remain, then 2 and finally 4
// lanes per row, which split the row's dot product (the late columns have the longest dots and the fewest rows).
__device__ __forceinline__ bool nw_chol(float *H, int n, int lane) {
    bool ok = true;
    for (int j = 0; j < n; j++) {
        const float *rj = H + j * (j + 1) / 2;
        const int rows = n - j;
        const float djj = rj[j];   // read before any lane overwrites the diagonal (the shuffle below is the rendez-vous)
        float piv;
        float *rown = nullptr;
        if (rows > 16) {
            int i0 = j + lane, i1 = j + lane + 32;
            float s0 = 0.f, s1 = 0.f;
            float *r0 = H + i0 * (i0 + 1) / 2, *r1 = H + i1 * (i1 + 1) / 2;
            if (i0 < n) {
                float a = r0[j], b = 0.f;
                int k = 0;
                for (; k + 1 < j; k += 2) { a -= r0[k] * rj[k]; b -= r0[k + 1] * rj[k + 1]; }
                if (k < j) a -= r0[k] * rj[k];
                s0 = a + b;
            }
            if (i1 < n) {
                s1 = r1[j];
                for (int k = 0; k < j; k++) s1 -= r1[k] * rj[k];
            }
            piv = __shfl_sync(AV_FULL, s0, 0);
            {   // modified Cholesky: a pivot lost to cancellation (fp32, |H_jj| >> the true pivot) is floored relative to the
                // diagonal it came from, so the direction stays a bounded descent direction and the line search still decreases the cost
                float fl = 2e-7f * fabsf(djj) + 1e-20f;
#ifdef AV_EMU_DEBUG
                if (!(piv > fl) && lane == 0) fprintf(stderr, "chol: pivot %d = %g (diag before %g)\n", j, piv, djj);
#endif
                if (!(piv > fl)) { ok = false; piv = fl; }
            }
            float rd = rsqrtf(piv);
            if (i0 < n) r0[j] = lane == 0 ? rd : s0 * rd;
            if (i1 < n) r1[j] = s1 * rd;
        } else {
            const int lpr = rows > 8 ? 2 : 4, sh = rows > 8 ? 1 : 2;   // lanes per row
            int i = j + (lane >> sh), part = lane & (lpr - 1);
            float s = 0.f;
            rown = H + i * (i + 1) / 2;
            if (i < n) {
                for (int k = part; k < j; k += lpr) s -= rown[k] * rj[k];
            }
            s += __shfl_xor_sync(AV_FULL, s, 1);
            if (lpr == 4) s += __shfl_xor_sync(AV_FULL, s, 2);
            if (i < n) s += rown[j];
            piv = __shfl_sync(AV_FULL, s, 0);
            {   // modified Cholesky: a pivot lost to cancellation (fp32, |H_jj| >> the true pivot) is floored relative to the
                // diagonal it came from, so the direction stays a bounded descent direction and the line search still decreases the cost
                float fl = 2e-7f * fabsf(djj) + 1e-20f;
#ifdef AV_EMU_DEBUG
                if (!(piv > fl) && lane == 0) fprintf(stderr, "chol: pivot %d = %g (diag before %g)\n", j, piv, djj);
#endif
                if (!(piv > fl)) { ok = false; piv = fl; }
            }
            float rd = rsqrtf(piv);
            if (i < n && part == 0) rown[j] = i == j ? rd : s * rd;
        }
        __syncwarp();
    }
    return ok;
}
// x = -(L L')^-1 g; lane holds rows lane and lane + 32 of the right-hand side / solution in registers
__device__ __forceinline__ void nw_solve(const float *H, int n, const float *g, int lane, float &x0, float &x1) {
    int i0 = lane, i1 = lane + 32;
    float s0 = i0 < n ? -g[i0] : 0.f, s1 = i1 < n ? -g[i1] : 0.f;
    for (int j = 0; j < n; j++) {
        float yj = __shfl_sync(AV_FULL, j < 32 ? s0 : s1, j & 31) * H[j * (j + 1) / 2 + j];
        if (lane == (j & 31)) { if (j < 32) s0 = yj; else s1 = yj; }
        if (i0 > j && i0 < n) s0 -= H[i0 * (i0 + 1) / 2 + j] * yj;
        if (i1 > j && i1 < n) s1 -= H[i1 * (i1 + 1) / 2 + j] * yj;
    }
    for (int j = n - 1; j >= 0; j--) {
        const float *rj = H + j * (j + 1) / 2;
        float xj = __shfl_sync(AV_FULL, j < 32 ? s0 : s1, j & 31) * rj[j];
        if (lane == (j & 31)) { if (j < 32) s0 = xj; else s1 = xj; }
        if (i0 < j) s0 -= rj[i0] * xj;
        if (i1 < j) s1 -= rj[i1] * xj;
    }
    x0 = s0; x1 = s1;
}

// The solve, as a state machine so that a kernel can run its phases in lockstep across warps (avsim_solve_kernel: one fetched
// instruction line then serves all the warps of the SM, like the stages of the substep kernel) -- or back to back for one warp
// (stage_newton below: forward kernel, fused step kernel, emulation harness).  In: scalar rows (S.sc_*), contact blocks
// (scratch), S.warm = previous qacc, S.qacc_smooth.  Out: S.acc (S.warm and S.qfrc_bias are clobbered) and the constraint forces
// S.sc_f / S.c_f of the final iterate (the noslip sweeps and the outputs read them).
//   init -> loop { grad (stop?) -> hess -> dir (stop?) -> search (stop?) } -> publish
struct Newton {
    float scale, fs, fa[6], fb[6], p0, p1, gn;   // registers that live across the phases of one solve
    int mycls, it;
    bool sfree;

    __device__ __forceinline__ void init(const DevModel &m, EnvS &S, float *scratch, int lane) {
        const int nv = m.nv, nsc = S.nsc, ncon = S.ncon;
        const int i0 = lane, i1 = lane + 32;
        it = 0; p0 = p1 = gn = 0.f;
        // scale of the stopping tests: 1 / trace(M) (MuJoCo: 1 / (meaninertia * nv))
        float trm = 0.f;
        for (int i = lane; i < nv; i += 32) {
            int t = m.dof_tree[i], dl = i - m.tree_dofadr[t];
            trm += S.M[t * AV_MTRI + av_mtri(dl, dl)];
        }
        scale = 1.0f / warp_sum(trm);
        mycls = lane < nsc ? (S.sc_key[lane] >> 24) : -1;
        // ---- start: the better of qacc_smooth (acc = 0) and the warm start (previous qacc), like mj_fwdConstraint
        for (int i = lane; i < AV_NVP; i += 32) S.acc[i] = i < nv ? S.warm[i] - S.qacc_smooth[i] : 0.f;
        __syncwarp();
        nw_jar_all(S, scratch, S.acc, lane);
        {
            float ga = 0.f;
            if (i0 < nv) ga += 0.5f * S.acc[i0] * nw_mrow(m, S, S.acc, i0);
            if (i1 < nv) ga += 0.5f * S.acc[i1] * nw_mrow(m, S, S.acc, i1);
            float cw = warp_sum(ga) + nw_cost(S, scratch, S.acc, false, lane), c0 = nw_cost(S, scratch, S.acc, true, lane);
            if (!(cw < c0)) {
                __syncwarp();
                for (int i = lane; i < AV_NVP; i += 32) S.acc[i] = 0.f;
                for (int c = 0; c < ncon; c++)
                    if (lane < 6) scratch[c * AV_CBLK + AV_CB_JAR + lane] = ldblk1(scratch + c * AV_CBLK + AV_CB_B + lane);
                __syncwarp();
            }
        }
    }

    // forces of the current iterate (lane = constraint block) and gradient g = M acc - J' f; true = stop here
    __device__ __forceinline__ bool grad(const DevModel &m, EnvS &S, float *scratch, int lane, int max_iter, float tol) {
        const int nv = m.nv, nsc = S.nsc, ncon = S.ncon;
        const int i0 = lane, i1 = lane + 32;
        float *g = S.qfrc_bias;
        fs = 0.f; sfree = false;
#pragma unroll
        for (int k = 0; k < 6; k++) fa[k] = fb[k] = 0.f;
        if (lane < nsc) fs = sc_force(S, lane, sc_jx(S, lane, S.acc) + S.sc_b[lane], sfree);
        for (int s = 0; s < 2; s++) {
            int c = lane + 32 * s;
            if (c < ncon && !((S.c_info[c] >> 20) & 1)) {
                ConeP cp;
                cone_load(scratch + c * AV_CBLK, cp);
                float y[6], f[6];
#pragma unroll
                for (int k = 0; k < 6; k++) y[k] = ldblk1(nw_jar(scratch, c) + k);
                cone_force(y, cp, f);
#pragma unroll
                for (int k = 0; k < 6; k++) { if (s == 0) fa[k] = f[k]; else fb[k] = f[k]; }
            }
        }
        if (i0 < nv) g[i0] = nw_mrow(m, S, S.acc, i0);
        if (i1 < nv) g[i1] = nw_mrow(m, S, S.acc, i1);
        __syncwarp();
        for (int cls = 0; cls < 3; cls++) {   // rows of one class touch distinct dofs: plain read-modify-write, deterministic
            if (mycls == cls) {
                g[S.sc_dof1[lane]] -= S.sc_c1[lane] * fs;
                if (S.sc_dof2[lane] >= 0) g[S.sc_dof2[lane]] -= S.sc_c2[lane] * fs;
            }
            __syncwarp();
        }
        for (int c = 0; c < ncon; c++) {
            if ((S.c_info[c] >> 20) & 1) continue;
            const int col = lane & 15, half = lane >> 4, tr = c_tr(S, c), dof = tr_dof(tr, col);
            const float *J = scratch + c * AV_CBLK + AV_CB_J + (3 * half) * AV_JW + col;
            float f[6];
#pragma unroll
            for (int k = 0; k < 6; k++) f[k] = __shfl_sync(AV_FULL, c < 32 ? fa[k] : fb[k], c & 31);
            float t = half ? (ldblk1(J) * f[3] + ldblk1(J + AV_JW) * f[4] + ldblk1(J + 2 * AV_JW) * f[5])
                           : (ldblk1(J) * f[0] + ldblk1(J + AV_JW) * f[1] + ldblk1(J + 2 * AV_JW) * f[2]);
            t += __shfl_xor_sync(AV_FULL, t, 16);
            if (half == 0 && dof >= 0) g[dof] -= t;
            __syncwarp();
        }
        gn = 0.f;
        if (i0 < nv) gn += g[i0] * g[i0];
        if (i1 < nv) gn += g[i1] * g[i1];
        gn = warp_sum(gn);
        return it >= max_iter || scale * sqrtf(gn) < tol;
    }

    // Hessian H = M + J' Hc J into the packed lower triangle S.H
    __device__ __forceinline__ void hess(const DevModel &m, EnvS &S, float *scratch, int lane) {
        const int nv = m.nv, ncon = S.ncon;
        float *H = S.H;
        // ---- Hessian H = M + J' Hc J
        const int nh = nv * (nv + 1) / 2;
        for (int i = lane; i < nh; i += 32) H[i] = 0.f;
        __syncwarp();
        for (int i = lane; i < nv; i += 32) {
            int t = m.dof_tree[i], d0 = m.tree_dofadr[t], dl = i - d0;
            for (int j = 0; j <= dl; j++) H[i * (i + 1) / 2 + d0 + j] = S.M[t * AV_MTRI + av_mtri(dl, j)];
        }
        __syncwarp();
        for (int cls = 0; cls < 3; cls++) {
            if (mycls == cls && sfree) {
                float D = __fdividef(1.0f, S.sc_R[lane]), c1 = S.sc_c1[lane], c2 = S.sc_c2[lane];
                int d1 = S.sc_dof1[lane], d2 = S.sc_dof2[lane];
                H[d1 * (d1 + 1) / 2 + d1] += D * c1 * c1;
                if (d2 >= 0) {
                    H[d2 * (d2 + 1) / 2 + d2] += D * c2 * c2;
                    H[nw_tri(d1, d2)] += D * c1 * c2;
                }
            }
            __syncwarp();
        }
        for (int c = 0; c < ncon; c++) {
            if ((S.c_info[c] >> 20) & 1) continue;
            const float *blk = scratch + c * AV_CBLK, *J = blk + AV_CB_J;
            ConeP cp;
            cone_load(blk, cp);
            float y[6], U[6], T_, e_;
#pragma unroll
            for (int k = 0; k < 6; k++) y[k] = ldblk1(nw_jar(scratch, c) + k);
            const int zone = cone_zone(y, cp, U, T_, e_);   // uniform: every lane evaluates the same contact
            if (zone == 0) continue;
            const int col = lane & 15, half = lane >> 4, tr = c_tr(S, c);
            float jc[6];
#pragma unroll
            for (int k = 0; k < 6; k++) jc[k] = ldblk1(J + k * AV_JW + col);
            if (zone == 1) {   // inside the cone: Hc = diag(1 / R), T = Hc J is a row scaling
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    float t0 = __fdividef(jc[r], cp.R[r]), t1 = __fdividef(jc[r + 3], cp.R[r + 3]);
                    S.stage[(3 * half + r) * AV_JW + col] = half ? t1 : t0;
                }
            } else {
                float Hc[21];
                cone_hess(y, cp, Hc);
#pragma unroll
                for (int r = 0; r < 3; r++) {   // T = Hc J: this lane's rows 3 half .. 3 half + 2 at its column
                    float t0 = 0.f, t1 = 0.f;
#pragma unroll
                    for (int l = 0; l < 6; l++) { t0 += Hc[TRI(r, l)] * jc[l]; t1 += Hc[TRI(r + 3, l)] * jc[l]; }
                    S.stage[(3 * half + r) * AV_JW + col] = half ? t1 : t0;
                }
            }
            __syncwarp();
            const int b1 = tr & 63, n1 = (tr >> 6) & 15, b2 = (tr >> 10) & 63, n2 = (tr >> 16) & 15, nl = n1 + n2, np = nl * (nl + 1) / 2;
            for (int e = lane; e < np; e += 32) {
                int ij = av_tri_ij[e], ii = ij >> 4, jj = ij & 15;
                int ci = ii < n1 ? ii : 8 + ii - n1, cj = jj < n1 ? jj : 8 + jj - n1;
                int di = ii < n1 ? b1 + ii : b2 + ii - n1, dj = jj < n1 ? b1 + jj : b2 + jj - n1;
                float s = 0.f;
#pragma unroll
                for (int k = 0; k < 6; k++) s += ldblk1(J + k * AV_JW + ci) * S.stage[k * AV_JW + cj];
                H[nw_tri(di, dj)] += s;
            }
            __syncwarp();
        }
    }

    // Cholesky, search direction p = -H^-1 g, Newton decrement; true = stop (the model predicts no further decrease)
    __device__ __forceinline__ bool dir(const DevModel &m, EnvS &S, int lane, float tol) {
        const int nv = m.nv;
        const int i0 = lane, i1 = lane + 32;
        float *g = S.qfrc_bias, *p = S.warm, *H = S.H;
        bool stop;
        if (!nw_chol(H, nv, lane)) S.status |= 16;
        nw_solve(H, nv, g, lane, p0, p1);
        if (i0 < nv) p[i0] = p0;
        if (i1 < nv) p[i1] = p1;
        float dec = -((i0 < nv ? g[i0] * p0 : 0.f) + (i1 < nv ? g[i1] * p1 : 0.f));
        dec = warp_sum(dec);   // Newton decrement: the decrease the quadratic model predicts is dec / 2
        __syncwarp();
        stop = !(0.5f * scale * dec >= 1e-6f * tol);   // also catches NaN / a non-descent direction
        return stop;
    }

    // exact line search along p and the update of acc / jar; true = stop (no descent left in fp32)
    __device__ __forceinline__ bool search(const DevModel &m, EnvS &S, float *scratch, int lane, int ls_iter) {
        const int nv = m.nv, nsc = S.nsc, ncon = S.ncon;
        const int i0 = lane, i1 = lane + 32;
        float *p = S.warm;
        bool stop = false;
        // ---- line search along p
        float pMp = 0.f, pMa = 0.f;
        if (i0 < nv) { float mp = nw_mrow(m, S, p, i0); pMp += p0 * mp; pMa += S.acc[i0] * mp; }
        if (i1 < nv) { float mp = nw_mrow(m, S, p, i1); pMp += p1 * mp; pMa += S.acc[i1] * mp; }
        pMp = warp_sum(pMp); pMa = warp_sum(pMa);
        float va[6], vb[6];
#pragma unroll
        for (int k = 0; k < 6; k++) va[k] = vb[k] = 0.f;
        for (int c = 0; c < ncon; c++) {
            if ((S.c_info[c] >> 20) & 1) continue;
            float res[6];
            nw_rows(scratch + c * AV_CBLK, c_tr(S, c), p, lane, res);
            if (lane == (c & 31)) {
#pragma unroll
                for (int k = 0; k < 6; k++) { if (c < 32) va[k] = res[k]; else vb[k] = res[k]; }
            }
        }
        LsCon La, Lb;
        ls_con_init(La, S, scratch, lane, va);
        ls_con_init(Lb, S, scratch, lane + 32, vb);
        float ys = 0.f, vs = 0.f, Rs = 1.f, los = 0.f, his = 0.f;
        if (lane < nsc) {
            ys = sc_jx(S, lane, S.acc) + S.sc_b[lane]; vs = sc_jx(S, lane, p);
            Rs = S.sc_R[lane]; los = S.sc_lo[lane]; his = S.sc_hi[lane];
        }
        float lo = 0.f, hi = -1.f, alpha = 0.f, d10 = 0.f;
        bool accepted = false;
        for (int k = 0; k <= ls_iter; k++) {
            float d1 = 0.f, d2 = 0.f;
            ls_con_eval(La, alpha, d1, d2);
            ls_con_eval(Lb, alpha, d1, d2);
            if (lane < nsc) {
                float y = ys + alpha * vs, f = -__fdividef(y, Rs);
                bool fr = f > los && f < his;
                f = fminf(fmaxf(f, los), his);
                d1 -= f * vs;
                if (fr) d2 += __fdividef(vs * vs, Rs);
            }
            d1 = warp_sum(d1) + pMa + alpha * pMp;
            d2 = warp_sum(d2) + pMp;
            if (k == 0) d10 = d1;
            else if (fabsf(d1) <= 1e-3f * fabsf(d10)) { accepted = true; break; }
            if (k == ls_iter) break;
            if (d1 < 0.f) lo = alpha; else hi = alpha;
            float next = alpha - __fdividef(d1, fmaxf(d2, 1e-30f));
            if (!(next > lo) || (hi >= 0.f && !(next < hi))) next = hi >= 0.f ? 0.5f * (lo + hi) : 2.f * alpha + 1e-6f;
            alpha = next;
        }
        if (!accepted) alpha = lo;   // the descending side of the bracket: the cost cannot increase
        if (!(d10 < 0.f) || !(alpha > 0.f)) stop = true;
        else {
            if (i0 < nv) S.acc[i0] += alpha * p0;
            if (i1 < nv) S.acc[i1] += alpha * p1;
            if (La.live) {
#pragma unroll
                for (int k = 0; k < 6; k++) nw_jar(scratch, lane)[k] += alpha * La.v[k];
            }
            if (Lb.live) {
#pragma unroll
                for (int k = 0; k < 6; k++) nw_jar(scratch, lane + 32)[k] += alpha * Lb.v[k];
            }
            __syncwarp();
            it++;
        }
        return stop;
    }

    // the forces of the last evaluated iterate become the solver's output; the Hessian (which overlays c_f / c_lam) is dead
    __device__ __forceinline__ void publish(EnvS &S, int lane, float &grad_out) {
        const int nsc = S.nsc, ncon = S.ncon;
        grad_out = scale * sqrtf(gn);
        if (lane < nsc) S.sc_f[lane] = fs;
        __syncwarp();
        if (lane < ncon) {   // the Hessian (which overlays c_f / c_lam) is dead from here on
#pragma unroll
            for (int k = 0; k < 6; k++) S.c_f[6 * lane + k] = fa[k];
            S.c_lam[lane] = 0.f;
        }
        if (lane + 32 < ncon) {
#pragma unroll
            for (int k = 0; k < 6; k++) S.c_f[6 * (lane + 32) + k] = fb[k];
            S.c_lam[lane + 32] = 0.f;
        }
        __syncwarp();
    }
};

__device__ AV_STAGE int stage_newton(const DevModel &m, EnvS &S, float *scratch, int lane, int max_iter, int ls_iter, float tol, float &grad_out, Prof &pf) {
    Newton nw;
    nw.init(m, S, scratch, lane);
    pf.mark(PF_NW_INIT, lane);
    for (;;) {
        bool stop = nw.grad(m, S, scratch, lane, max_iter, tol);
        pf.mark(PF_NW_GRAD, lane);
        if (stop) break;
        nw.hess(m, S, scratch, lane);
        pf.mark(PF_NW_HESS, lane);
        stop = nw.dir(m, S, lane, tol);
        pf.mark(PF_NW_CHOL, lane);
        if (stop) break;
        stop = nw.search(m, S, scratch, lane, ls_iter);
        pf.mark(PF_NW_LS, lane);
        if (stop) break;
    }
    nw.publish(S, lane, grad_out);
    return nw.it;
}
