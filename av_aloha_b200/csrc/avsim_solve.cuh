// avsim_solve.cuh -- K5 (contact rows) and K6 (block projected Gauss-Seidel on the dual + noslip), one warp per env.
//
// Replaces mj_makeConstraint's contact rows and the constraint solve that `physics.step` (reference env.py:218) runs
// inside MuJoCo [third-party; semantics per SURVEY.md Appendix A]: elliptic cones, impratio, soft-constraint
// reference acceleration, warm start from the previous qacc, then `noslip_iterations` sweeps on the unregularised
// friction rows.  Same statement as oracle/avsim_oracle.c stage_constraint_rows / stage_solve, fp32.
//
// Data layout.  Per contact one 592-byte block in a per-environment global scratch (read-only during the sweeps, so it
// lives in L1): [AR 21 | Lc 15 | b 6 | Rn Rf Rt Rr | mu0 mu1 mu2 | 1/mu0 1/mu1 1/mu2] = 13 float4, then J[6][16]
// (8 dof columns of tree 1 | 8 of tree 2).  Cholesky factors keep the RECIPROCAL of their diagonal, so the triangular
// solves of the sweeps are multiply-only (one MUFU.RSQ per pivot when factoring, none when solving).  The mutable parts -- the force f[6] per contact and the constraint acceleration acc[nv] -- stay in
// shared memory.  M^-1 J^T is NOT stored: an update applies  acc += Minv_tree (J^T df)  with the 8x8 block inverses
// that are already in shared memory (4 shuffles + 4 FMAs per lane), halving the block traffic.
// Everything below indexes its register arrays with compile-time constants only (all loops fully unrolled at the
// fixed sizes 6 / 5; a condim-3 contact carries three dead rows with J = 0, b = 0 and a unit diagonal), so nothing
// spills to local memory.  Lane roles: lane = (half = lane >> 4 -> rows 3*half..3*half+2, col = lane & 15).
#pragma once

#define AV_CB_AR 0
#define AV_CB_LC 21
#define AV_CB_B 36
#define AV_CB_PAR 42
#define AV_CB_J 52
// AV_CB_GEO = 148: contact geometry; AV_CBLK (avsim_dev.h) = 164 floats = 656 bytes

#define TRI(i, j) ((i) >= (j) ? (i) * ((i) + 1) / 2 + (j) : (j) * ((j) + 1) / 2 + (i))

// packed tree descriptor of a contact: dofadr1 | n1 << 6 | dofadr2 << 10 | n2 << 16 | tree1 << 20 | tree2 << 23
__device__ __forceinline__ int tr_pack(const DevModel &m, int t1, int t2) {
    int v = 0;
    if (t1 >= 0) v |= m.tree_dofadr[t1] | (m.tree_dofnum[t1] << 6) | (t1 << 20);
    if (t2 >= 0) v |= (m.tree_dofadr[t2] << 10) | (m.tree_dofnum[t2] << 16) | (t2 << 23);
    return v;
}
__device__ __forceinline__ int tr_dof(int tr, int col) {   // dof of this lane's column, -1 when the column is empty
    int sh = (col & 8) ? 10 : 0, base = (tr >> sh) & 63, n = (tr >> (sh + 6)) & 15, dl = col & 7;
    return dl < n ? base + dl : -1;
}
__device__ __forceinline__ int tr_tree(int tr, int col) { return (tr >> ((col & 8) ? 23 : 20)) & 7; }
// the same descriptor rebuilt from the two tree ids kept in c_info (7 = none) and the per-tree table in shared memory
__device__ __forceinline__ int c_tr(const EnvS &S, int c) {
    int info = S.c_info[c], t1 = (info >> 21) & 7, t2 = (info >> 24) & 7;
    return (t1 < 7 ? S.tree_pk[t1] | (t1 << 20) : 0) | (t2 < 7 ? (S.tree_pk[t2] << 10) | (t2 << 23) : 0);
}

// sum of three per-lane values over the 16 lanes of a half-warp; every lane of the half gets the sums
__device__ __forceinline__ void half_sum3(float &r0, float &r1, float &r2) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        r0 += __shfl_xor_sync(AV_FULL, r0, o);
        r1 += __shfl_xor_sync(AV_FULL, r1, o);
        r2 += __shfl_xor_sync(AV_FULL, r2, o);
    }
}
// J (this lane's three rows j0..j2 at its column) times a joint-space vector x (already gathered per column):
// all six row results in every lane
__device__ __forceinline__ void block_rows(float j0, float j1, float j2, float x, int half, float (&res)[6]) {
    float r0 = j0 * x, r1 = j1 * x, r2 = j2 * x;
    half_sum3(r0, r1, r2);
    float o0 = __shfl_xor_sync(AV_FULL, r0, 16), o1 = __shfl_xor_sync(AV_FULL, r1, 16), o2 = __shfl_xor_sync(AV_FULL, r2, 16);
    res[0] = half ? o0 : r0; res[1] = half ? o1 : r1; res[2] = half ? o2 : r2;
    res[3] = half ? r0 : o0; res[4] = half ? r1 : o1; res[5] = half ? r2 : o2;
}
// acc += Minv_tree (J^T df) for this contact
__device__ __forceinline__ void block_apply(EnvS &S, float j0, float j1, float j2, const float (&df)[6], int tr, int lane) {
    int col = lane & 15, half = lane >> 4, dl = col & 7;
    float t = half ? (j0 * df[3] + j1 * df[4] + j2 * df[5]) : (j0 * df[0] + j1 * df[1] + j2 * df[2]);
    t += __shfl_xor_sync(AV_FULL, t, 16);            // (J^T df)[col], both halves
    const float *Mi = S.Minv + tr_tree(tr, col) * (AV_TD * AV_TD) + dl * AV_TD + 4 * half;
    float da = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) da += Mi[k] * __shfl_sync(AV_FULL, t, (col & 8) + 4 * half + k, 16);
    da += __shfl_xor_sync(AV_FULL, da, 16);
    int dof = tr_dof(tr, col);
    if (half == 0 && dof >= 0) S.acc[dof] += da;
}

struct CBlk {   // the uniform (per-contact) part of a block, in registers
    float AR[21], Lc[15], b[6], R[6], mu[5], imu[5];
};
__device__ __forceinline__ void cblk_load(const float *blk, CBlk &c) {
    const float4 *p = reinterpret_cast<const float4 *>(blk);
    float v[52];
#pragma unroll
    for (int i = 0; i < 13; i++) {
        float4 q = ldblk4(p + i);
        v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w;
    }
#pragma unroll
    for (int i = 0; i < 21; i++) c.AR[i] = v[AV_CB_AR + i];
#pragma unroll
    for (int i = 0; i < 15; i++) c.Lc[i] = v[AV_CB_LC + i];
#pragma unroll
    for (int i = 0; i < 6; i++) c.b[i] = v[AV_CB_B + i];
    c.R[0] = v[AV_CB_PAR]; c.R[1] = c.R[2] = v[AV_CB_PAR + 1]; c.R[3] = v[AV_CB_PAR + 2]; c.R[4] = c.R[5] = v[AV_CB_PAR + 3];
    c.mu[0] = c.mu[1] = v[AV_CB_PAR + 4]; c.mu[2] = v[AV_CB_PAR + 5]; c.mu[3] = c.mu[4] = v[AV_CB_PAR + 6];
    c.imu[0] = c.imu[1] = v[AV_CB_PAR + 7]; c.imu[2] = v[AV_CB_PAR + 8]; c.imu[3] = c.imu[4] = v[AV_CB_PAR + 9];
}

// ---- 5x5 helpers on packed lower triangles, compile-time indices only.  L holds 1/L_ii on its diagonal.
__device__ __forceinline__ void chol5(const float (&A)[15], float lam, float (&L)[15]) {
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            float s = A[TRI(i, j)] + (i == j ? lam : 0.f);
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[TRI(i, k)] * L[TRI(j, k)];
            L[TRI(i, j)] = (i == j) ? rsqrtf(fmaxf(s, AV_MINVAL)) : s * L[TRI(j, j)];
        }
}
// solve L L^T x = -b
__device__ __forceinline__ void trisolve5(const float (&L)[15], const float (&b)[5], float (&x)[5]) {
#pragma unroll
    for (int i = 0; i < 5; i++) {
        float s = -b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[TRI(i, k)] * x[k];
        x[i] = s * L[TRI(i, i)];
    }
#pragma unroll
    for (int i = 4; i >= 0; i--) {
        float s = x[i];
#pragma unroll
        for (int k = i + 1; k < 5; k++) s -= L[TRI(k, i)] * x[k];
        x[i] = s * L[TRI(i, i)];
    }
}
// minimise 0.5 y'Ay + y'b  s.t.  sum (y_i/mu_i)^2 <= r^2  (A packed lower 5x5).  L0 = Cholesky factor of A when
// have_L0, otherwise it is computed here.  `lam` carries the multiplier of the previous sweep for this contact:
// the local problems barely change from sweep to sweep, so the Newton iteration on |z(lam)| = r restarts next to its root.
__device__ __forceinline__ void qcqp5(const float (&A)[15], const float (&b)[5], const float (&mu)[5], const float (&imu)[5],
                                      float r, const float (&L0)[15], bool have_L0, float &lam, float (&y)[5]) {
    float L[15];
    if (have_L0) {
#pragma unroll
        for (int i = 0; i < 15; i++) L[i] = L0[i];
    } else
        chol5(A, 0.f, L);
    trisolve5(L, b, y);
    float zz = 0.f;
#pragma unroll
    for (int i = 0; i < 5; i++) zz += (y[i] * imu[i]) * (y[i] * imu[i]);
    if (zz <= r * r) { lam = 0.f; return; }
    // scaled problem z = y / mu; Newton on the multiplier of |z| = r
    float As[15], bs[5], z[5], w[5], mz[5];
#pragma unroll
    for (int i = 0; i < 5; i++) {
        bs[i] = b[i] * mu[i];
#pragma unroll
        for (int j = 0; j <= i; j++) As[TRI(i, j)] = A[TRI(i, j)] * mu[i] * mu[j];
    }
    float ir = 1.0f / r;
    for (int it = 0; it < 12; it++) {
        chol5(As, lam, L);
        trisolve5(L, bs, z);
        zz = 0.f;
#pragma unroll
        for (int i = 0; i < 5; i++) { zz += z[i] * z[i]; mz[i] = -z[i]; }
        if (fabsf(zz - r * r) < 1e-6f * fmaxf(1e-12f, r * r)) break;
        trisolve5(L, mz, w);
        float zw = 0.f;
#pragma unroll
        for (int i = 0; i < 5; i++) zw += z[i] * w[i];
        float nz = sqrtf(zz);
        lam = fmaxf(0.f, lam + (nz - r) * ir * zz / fmaxf(zw, AV_MINVAL));
    }
    if (zz > r * r) {
        float s = r * rsqrtf(zz);
#pragma unroll
        for (int i = 0; i < 5; i++) z[i] *= s;
    }
#pragma unroll
    for (int i = 0; i < 5; i++) y[i] = z[i] * mu[i];
}

// ------------------------------------------------------------------ K5: contact rows, one contact at a time
__device__ __forceinline__ int contact_key(const EnvS &S, int c) {
    int pair = S.c_info[c] & 0xffff, ord = 0;
    for (int k = c - 1; k >= 0 && (S.c_info[k] & 0xffff) == pair; k--) ord++;
    return (3 << 24) | pair | (ord << 16);
}

__device__ AV_STAGE void stage_rows_contact(const DevModel &m, EnvS &S, float *scratch, int lane, const FCache &fc) {
    int col = lane & 15, half = lane >> 4;
    for (int c = 0; c < S.ncon; c++) {
        int info = S.c_info[c], g1 = info & 0xff, g2 = (info >> 8) & 0xff, dim = (info >> 16) & 0xf;
        if ((info >> 20) & 1) continue;   // detected but outside margin - gap (pin geoms): no rows
        int b1 = m.geom_body[g1], b2 = m.geom_body[g2];
        int t1 = m.body_tree[b1], t2 = m.body_tree[b2];
        if (t1 < 0) { t1 = t2; t2 = -1; }   // keep the first slot occupied; signs are handled per body below
        if (t2 == t1) t2 = -1;
        int tr = tr_pack(m, t1, t2);
        int t = col < 8 ? t1 : t2, dl = col & 7, dof = tr_dof(tr, col);
        float j0 = 0.f, j1 = 0.f, j2 = 0.f;
        const float *geo = scratch + c * AV_CBLK + AV_CB_GEO;
        V3 p = ld3(geo);
        V3 fn = ld3(geo + 3), ft1 = ld3(geo + 6), ft2 = ld3(geo + 9);
        if (dof >= 0) {
            float sgn = 0.f;
            if (m.body_tree[b2] == t && ((m.body_treemask[b2] >> dl) & 1)) sgn += 1.f;
            if (m.body_tree[b1] == t && ((m.body_treemask[b1] >> dl) & 1)) sgn -= 1.f;
            if (sgn != 0.f) {
                S6 cd = ld6(S.cdof + 6 * dof);
                V3 v = half == 0 ? cd.l + cross(cd.a, p - ld3(S.torig + 3 * t)) : cd.a;
                v = v * sgn;
                j0 = dot(fn, v); j1 = dot(ft1, v); j2 = dot(ft2, v);
            }
        }
        if (half == 1 && dim < 6) { j0 = j1 = j2 = 0.f; }   // condim 3: no torsional / rolling rows
        float *Js = S.stage, *MJs = S.stage + 6 * AV_JW;
        Js[(3 * half + 0) * AV_JW + col] = j0; Js[(3 * half + 1) * AV_JW + col] = j1; Js[(3 * half + 2) * AV_JW + col] = j2;
        __syncwarp();
        // M^-1 J^T rows (only needed to form A = J M^-1 J^T here)
        float m0 = 0.f, m1 = 0.f, m2 = 0.f;
        if (dof >= 0) {
            const float *Mi = S.Minv + t * AV_TD * AV_TD + dl;  // column dl (symmetric)
            int cb = col & 8;
#pragma unroll
            for (int k = 0; k < AV_TD; k++) {
                float mv = Mi[k * AV_TD];
                m0 += Js[(3 * half + 0) * AV_JW + cb + k] * mv;
                m1 += Js[(3 * half + 1) * AV_JW + cb + k] * mv;
                m2 += Js[(3 * half + 2) * AV_JW + cb + k] * mv;
            }
        }
        MJs[(3 * half + 0) * AV_JW + col] = m0; MJs[(3 * half + 1) * AV_JW + col] = m1; MJs[(3 * half + 2) * AV_JW + col] = m2;
        float *blk = scratch + c * AV_CBLK;
        blk[AV_CB_J + (3 * half + 0) * AV_JW + col] = j0; blk[AV_CB_J + (3 * half + 1) * AV_JW + col] = j1;
        blk[AV_CB_J + (3 * half + 2) * AV_JW + col] = j2;
        // velocities / smooth accelerations / warm-start accelerations along the rows
        float xv = dof >= 0 ? S.qvel[dof] : 0.f, xa = dof >= 0 ? S.qacc_smooth[dof] : 0.f, xw = dof >= 0 ? S.warm[dof] : 0.f;
        float vv[6], aa[6], ww[6];
        block_rows(j0, j1, j2, xv, half, vv);
        block_rows(j0, j1, j2, xa, half, aa);
        if (fc.mode == 1) block_rows(j0, j1, j2, xw, half, ww);   // J qacc_warmstart: only the MuJoCo-style warm start needs it
        else {
#pragma unroll
            for (int k = 0; k < 6; k++) ww[k] = 0.f;
        }
        __syncwarp();
        // impedance, regularisation, reference acceleration (uniform)
        float K, B, imp, solref[2], solimp[5];
#pragma unroll
        for (int k = 0; k < 2; k++) solref[k] = 0.5f * (m.geom_solref[2 * g1 + k] + m.geom_solref[2 * g2 + k]);
#pragma unroll
        for (int k = 0; k < 5; k++) solimp[k] = 0.5f * (m.geom_solimp[5 * g1 + k] + m.geom_solimp[5 * g2 + k]);
        float dist = geo[12];
        kbi(m, solref, solimp, dist, K, B, imp);
        float Rn = fmaxf(AV_MINVAL, (1.f - imp) / imp * (m.body_invweight0[2 * b1] + m.body_invweight0[2 * b2]));
        float mu0 = geo[13], mu1 = geo[14], mu2 = geo[15];
        float Rf = Rn / m.impratio;
        float R[6] = {Rn, Rf, Rf, Rf * mu0 * mu0 / (mu1 * mu1), Rf * mu0 * mu0 / (mu2 * mu2), Rf * mu0 * mu0 / (mu2 * mu2)};
        if (lane < 6) {
            // row = lane; select the row's values without dynamic register indexing
            float v = 0.f, a = 0.f, w = 0.f, Rr = 1.f;
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (k == lane) { v = vv[k]; a = aa[k]; w = ww[k]; Rr = R[k]; }
            float aref = -B * v - (lane == 0 ? K * imp * dist : 0.f);
            bool live = lane < dim;
            blk[AV_CB_B + lane] = live ? a - aref : 0.f;
            S.c_f[6 * c + lane] = live ? -(w - aref) / Rr : 0.f;
        }
        if (lane == 0) {
            blk[AV_CB_PAR] = R[0]; blk[AV_CB_PAR + 1] = R[1]; blk[AV_CB_PAR + 2] = R[3]; blk[AV_CB_PAR + 3] = R[4];
            blk[AV_CB_PAR + 4] = mu0; blk[AV_CB_PAR + 5] = mu1; blk[AV_CB_PAR + 6] = mu2;
            blk[AV_CB_PAR + 7] = 1.0f / mu0; blk[AV_CB_PAR + 8] = 1.0f / mu1; blk[AV_CB_PAR + 9] = 1.0f / mu2;
            S.c_info[c] = info | ((t1 < 0 ? 7 : t1) << 21) | ((t2 < 0 ? 7 : t2) << 24);
            S.c_lam[c] = 0.f;
        }
        // AR = J MinvJT^T + R, packed lower triangle (21 entries), lane = entry
        if (lane < 21) {
            int i = 0;
            while ((i + 1) * (i + 2) / 2 <= lane) i++;
            int j = lane - i * (i + 1) / 2;
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < AV_JW; k++) s += Js[i * AV_JW + k] * MJs[j * AV_JW + k];
            if (i == j) {
                float Rr = 1.f;   // dead rows get a unit diagonal
#pragma unroll
                for (int k = 0; k < 6; k++)
                    if (k == i && k < dim) Rr = R[k];
                s += Rr;
            }
            blk[AV_CB_AR + lane] = s;
        }
        __syncwarp();
    }
    // lane = contact: Cholesky factor of the regularised friction block (rows 1..5), packed lower (15);
    // projection of the warm-start force onto the cone
    for (int c = lane; c < S.ncon; c += 32) {
        int info = S.c_info[c];
        if ((info >> 20) & 1) continue;
        float *blk = scratch + c * AV_CBLK;
        float A[15], Lc[15];
#pragma unroll
        for (int i = 0; i < 5; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) A[TRI(i, j)] = blk[AV_CB_AR + TRI(i + 1, j + 1)];
        if (fc.mode == 3) {   // Newton solver: only the noslip sweeps use the factor, and they drop the regulariser of the live rows
            int dim = (info >> 16) & 0xf;
            float Rf = blk[AV_CB_PAR + 1], Rt = blk[AV_CB_PAR + 2], Rr = blk[AV_CB_PAR + 3];
            A[TRI(0, 0)] -= Rf; A[TRI(1, 1)] -= Rf;
            if (dim > 3) { A[TRI(2, 2)] -= Rt; A[TRI(3, 3)] -= Rr; A[TRI(4, 4)] -= Rr; }
        }
        chol5(A, 0.f, Lc);
#pragma unroll
        for (int k = 0; k < 15; k++) blk[AV_CB_LC + k] = Lc[k];
        if (fc.mode == 3) continue;   // the primal solver publishes its own forces
        float *f = S.c_f + 6 * c;
        if (fc.mode == 2) {   // warm start from the force this contact carried in the previous solve (0 when it is new)
            int key = contact_key(S, c), nc = fc.n[0], hit = -1;
            for (int k = 0; k < nc; k++)
                if (fc.key[k] == key) { hit = k; break; }
#pragma unroll
            for (int k = 0; k < 6; k++) f[k] = hit >= 0 ? fc.val[6 * hit + k] : 0.f;
        }
        if (f[0] <= 0.f) {
#pragma unroll
            for (int k = 0; k < 6; k++) f[k] = 0.f;
        } else {
            float mu0 = blk[AV_CB_GEO + 13], mu1 = blk[AV_CB_GEO + 14], mu2 = blk[AV_CB_GEO + 15];
            float mu[5] = {mu0, mu0, mu1, mu2, mu2};
            float s = 0.f;
#pragma unroll
            for (int k = 1; k < 6; k++) s += (f[k] / mu[k - 1]) * (f[k] / mu[k - 1]);
            if (s > f[0] * f[0]) {
                float sc = f[0] * rsqrtf(s);
#pragma unroll
                for (int k = 1; k < 6; k++) f[k] *= sc;
            }
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------ K6: block projected Gauss-Seidel + noslip
// One block update.  Regularised sweep (noslip = false): (a) ray update of the whole 6-vector, (b) friction rows on the
// ellipsoid of radius f_n.  Noslip sweep: friction rows only, unregularised A, normal force fixed.  Both modes share ONE
// instance of the QCQP code: the sweep loop is the hottest code of the kernel and has to stay inside the SM's
// instruction cache (ncu showed 'no_instruction' as the top stall when each mode carried its own copy).
__device__ __forceinline__ void contact_block_update(const CBlk &cb, int dim, bool noslip, bool lc_noslip, float (&res)[6], const float (&old)[6],
                                                     float &lam, float (&f)[6]) {
#pragma unroll
    for (int k = 0; k < 6; k++) { res[k] += cb.b[k] + (noslip ? 0.f : cb.R[k] * old[k]); f[k] = old[k]; }
    if (!noslip) {
        if (old[0] < AV_MINVAL) {
            f[0] = fmaxf(0.f, old[0] - __fdividef(res[0], cb.AR[0]));
#pragma unroll
            for (int k = 1; k < 6; k++) f[k] = 0.f;
        } else {
            float vAv = 0.f, vr = 0.f;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                vr += old[k] * res[k];
#pragma unroll
                for (int l = 0; l < 6; l++) vAv += old[k] * cb.AR[TRI(k, l)] * old[l];
            }
            if (vAv > AV_MINVAL) {
                float x = -__fdividef(vr, vAv);
                if (old[0] + x * old[0] < 0.f) x = -1.f;
#pragma unroll
                for (int k = 0; k < 6; k++) f[k] = old[k] + x * old[k];
            }
        }
    }
    if (f[0] < AV_MINVAL) {
#pragma unroll
        for (int k = 1; k < 6; k++) f[k] = 0.f;
        return;
    }
    float Ac[15], bc[5], y[5];
#pragma unroll
    for (int k = 1; k < 6; k++) {
        float s = res[k] + (noslip ? 0.f : cb.AR[TRI(k, 0)] * (f[0] - old[0]));
#pragma unroll
        for (int l = 1; l < 6; l++) {
            float a = cb.AR[TRI(k, l)] - ((noslip && k == l && k < dim) ? cb.R[k] : 0.f);
            if (l <= k) Ac[TRI(k - 1, l - 1)] = a;
            s -= a * old[l];
        }
        bc[k - 1] = s;
    }
    qcqp5(Ac, bc, cb.mu, cb.imu, f[0], cb.Lc, noslip == lc_noslip, lam, y);   // Lc factors the block of exactly one of the two modes
#pragma unroll
    for (int k = 1; k < 6; k++) f[k] = y[k - 1];
}

// one Gauss-Seidel sweep over the scalar rows and the contact blocks (the hot loop; its own function so that its code
// is contiguous and small)
__device__ __noinline__ void solve_sweep(const DevModel &m, EnvS &S, const float *scratch, int lane, bool noslip, bool lc_noslip = false) {
    int col = lane & 15, half = lane >> 4;
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
    // contacts, two at a time: the schedule (solve_schedule) pairs contacts that touch disjoint kinematic trees; such
    // updates commute, so each half-warp (16 lanes = the 16 columns of a contact Jacobian) runs its own contact's
    // update concurrently -- the block update itself is uniform per contact and used to be replicated on all 32 lanes
    for (int sl = 0; sl < S.nslot; sl++) {
        const int pr = S.c_slot[sl], ca = pr & 0xff, cb_ = (pr >> 8) & 0xff;
        const bool valid = !half || cb_ != 0xff;
        const int c = (half && cb_ != 0xff) ? cb_ : ca;        // an idle second half mirrors the first (results discarded)
        const int info = S.c_info[c], dim = (info >> 16) & 0xf;
        const float *blk = scratch + c * AV_CBLK;
        const int tr = c_tr(S, c), dof = tr_dof(tr, col);
        const float *J = blk + AV_CB_J;
        float jc[6], res[6], old[6], f[6], df[6];
#pragma unroll
        for (int k = 0; k < 6; k++) jc[k] = ldblk1(J + k * AV_JW + col);
        const float x = dof >= 0 ? S.acc[dof] : 0.f;
#pragma unroll
        for (int k = 0; k < 6; k++) res[k] = jc[k] * x;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < 6; k++) res[k] += __shfl_xor_sync(AV_FULL, res[k], o);
        CBlk cb;
        cblk_load(blk, cb);
#pragma unroll
        for (int k = 0; k < 6; k++) old[k] = S.c_f[6 * c + k];
        float lam = S.c_lam[c];
        contact_block_update(cb, dim, noslip, lc_noslip, res, old, lam, f);
#pragma unroll
        for (int k = 0; k < 6; k++) df[k] = f[k] - old[k];
        __syncwarp();
        // acc += Minv_tree (J^T df): this lane's column of J^T df, then the 8x8 block inverse of the column's tree
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 6; k++) t += jc[k] * df[k];
        const float *Mi = S.Minv + tr_tree(tr, col) * (AV_TD * AV_TD) + (col & 7) * AV_TD;
        float da = 0.f;
#pragma unroll
        for (int k = 0; k < AV_TD; k++) da += Mi[k] * __shfl_sync(AV_FULL, t, (col & 8) + k, 16);
        if (valid && dof >= 0) S.acc[dof] += da;
        if (valid && col < 6) {
            float fv = 0.f;
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (k == col) fv = f[k];
            S.c_f[6 * c + col] = fv;
        }
        if (valid && col == 6) S.c_lam[c] = lam;
        __syncwarp();
    }
}

// Half-warp schedule of the sweep: every not-yet-scheduled contact a is paired with the next later contact b whose
// kinematic trees are disjoint from a's (b = 0xff when there is none).  Same greedy rule, same order as the oracle.
__device__ inline void solve_schedule(EnvS &S, int lane) {
    auto tmask = [&](int c) {
        if (c >= S.ncon || ((S.c_info[c] >> 20) & 1)) return -1;          // no rows: never scheduled
        int tr = c_tr(S, c);
        return (((tr >> 6) & 15) ? 1 << ((tr >> 20) & 7) : 0) | (((tr >> 16) & 15) ? 1 << ((tr >> 23) & 7) : 0);
    };
    const int m0 = tmask(lane), m1 = tmask(lane + 32);
    unsigned u0 = __ballot_sync(AV_FULL, m0 < 0), u1 = __ballot_sync(AV_FULL, m1 < 0);   // "used" bit sets
    int ns = 0;
    for (int a = 0; a < S.ncon; a++) {
        if ((a < 32 ? u0 >> a : u1 >> (a - 32)) & 1u) continue;
        if (a < 32) u0 |= 1u << a; else u1 |= 1u << (a - 32);
        const int ma = __shfl_sync(AV_FULL, a < 32 ? m0 : m1, a & 31);
        unsigned later0 = a < 31 ? ~((2u << a) - 1u) : 0u, later1 = a < 32 ? 0xffffffffu : (a < 63 ? ~((2u << (a - 32)) - 1u) : 0u);
        unsigned b0 = __ballot_sync(AV_FULL, m0 >= 0 && !(m0 & ma)) & ~u0 & later0;
        unsigned b1 = __ballot_sync(AV_FULL, m1 >= 0 && !(m1 & ma)) & ~u1 & later1;
        int b = b0 ? __ffs(b0) - 1 : (b1 ? 32 + __ffs(b1) - 1 : 0xff);
        if (b != 0xff) { if (b < 32) u0 |= 1u << b; else u1 |= 1u << (b - 32); }
        if (lane == 0) S.c_slot[ns] = a | (b << 8);
        ns++;
    }
    if (lane == 0) S.nslot = ns;
    __syncwarp();
}

__device__ AV_STAGE void stage_solve_begin(const DevModel &m, EnvS &S, float *scratch, int lane, int warm_mode) {
    int col = lane & 15, half = lane >> 4;
    if (warm_mode == 3) {   // after the primal (Newton) solve: S.acc IS the solution and S.sc_f / S.c_f its forces.  The noslip
        solve_schedule(S, lane);   // sweeps add M^-1 J' df on top of it.  Rebuilding acc as M^-1 J' f instead (what a dual solver
        return;                    // has to do) would push the solve's residual g = M acc - J' f through M^-1 -- the UNCONSTRAINED
    }                              // inverse: 0.25 rad/s^2 on a finger dof from a scaled residual of 2.6e-5 (profiles/r2_newton_parity.txt)
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
        const float *J = scratch + c * AV_CBLK + AV_CB_J;
        float j0 = J[(3 * half) * AV_JW + col], j1 = J[(3 * half + 1) * AV_JW + col], j2 = J[(3 * half + 2) * AV_JW + col];
        float df[6];
#pragma unroll
        for (int k = 0; k < 6; k++) df[k] = S.c_f[6 * c + k];
        block_apply(S, j0, j1, j2, df, c_tr(S, c), lane);
        __syncwarp();
    }
    float cost = 0.f;
    if (warm_mode != 2) {   // dual cost of the warm start (decides whether MuJoCo-style warm start is kept)
        for (int r = 0; r < S.nsc; r++) {
            float f = S.sc_f[r];
            float ja = S.sc_c1[r] * S.acc[S.sc_dof1[r]] + (S.sc_dof2[r] >= 0 ? S.sc_c2[r] * S.acc[S.sc_dof2[r]] : 0.f);
            cost += f * (0.5f * (ja + S.sc_R[r] * f) + S.sc_b[r]);
        }
        for (int c = 0; c < S.ncon; c++) {
            if ((S.c_info[c] >> 20) & 1) continue;
            const float *blk = scratch + c * AV_CBLK;
            int tr = c_tr(S, c), dof = tr_dof(tr, col);
            const float *J = blk + AV_CB_J;
            float j0 = J[(3 * half) * AV_JW + col], j1 = J[(3 * half + 1) * AV_JW + col], j2 = J[(3 * half + 2) * AV_JW + col];
            float res[6];
            block_rows(j0, j1, j2, dof >= 0 ? S.acc[dof] : 0.f, half, res);
            // R (rows 0 | 1,2 | 3 | 4,5) and b straight from the block: the full CBlk is not needed here
            float Rn = blk[AV_CB_PAR], Rf = blk[AV_CB_PAR + 1], Rt = blk[AV_CB_PAR + 2], Rr = blk[AV_CB_PAR + 3];
            float Rk[6] = {Rn, Rf, Rf, Rt, Rr, Rr};
#pragma unroll
            for (int k = 0; k < 6; k++) {
                float f = S.c_f[6 * c + k];   // dead rows: f = 0
                cost += f * (0.5f * (res[k] + Rk[k] * f) + blk[AV_CB_B + k]);
            }
        }
        __syncwarp();
    }
    // MuJoCo keeps its warm start only if it beats f = 0; the force cache (mode 2) is kept unconditionally
    if (warm_mode != 2 && cost >= 0.f) {
        for (int i = lane; i < AV_NVP; i += 32) S.acc[i] = 0.f;
        for (int i = lane; i < AV_NSC; i += 32) S.sc_f[i] = 0.f;
        for (int i = lane; i < AV_NCON * 6; i += 32) S.c_f[i] = 0.f;
        __syncwarp();
    }
    solve_schedule(S, lane);
}
// after the last sweep: remember every constraint's force under its identity key for the next solve's warm start
__device__ inline void stage_cache_store(const DevModel &m, EnvS &S, int lane, const FCache &fc) {
    for (int c = lane; c < S.ncon; c += 32) {
        fc.key[c] = contact_key(S, c);
        bool live = !((S.c_info[c] >> 20) & 1);
#pragma unroll
        for (int k = 0; k < 6; k++) fc.val[6 * c + k] = live ? S.c_f[6 * c + k] : 0.f;
    }
    for (int r = lane; r < S.nsc; r += 32) { fc.key[AV_NCON + r] = S.sc_key[r]; fc.val[AV_NCON * 6 + r] = S.sc_f[r]; }
    if (lane == 0) { fc.n[0] = S.ncon; fc.n[1] = S.nsc; }
    __syncwarp();
}
__device__ inline void stage_solve(const DevModel &m, EnvS &S, float *scratch, int lane, int iters, int noslip_iters, int warm_mode, bool lc_noslip = false) {
    stage_solve_begin(m, S, scratch, lane, warm_mode);
    for (int it = 0; it < iters + noslip_iters; it++) solve_sweep(m, S, scratch, lane, it >= iters, lc_noslip);
}
