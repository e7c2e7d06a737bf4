// avsim_ik.cuh -- batched IK controllers K11 (DiffIK) and K12 (GradIK), one thread per (q, target) problem.
//
// Restates reference data_collection_scripts/kinematics.py:7-52 (product-of-exponentials FK, space Jacobian),
// diff_ik.py:51-85 (damped least squares + null-space term, clipped Euler integration),
// grad_ik.py:8-99,176-218 (finite-difference gradient descent with secant line step) and the
// transform_utils.py helpers they call (exp2mat 239-261, adjoint 289-301, angular_error 183-194,
// quat2mat 52-79 incl. its float32 round trip, limit_pose 263-287, mat2quat 9-49, quat2axisangle 82-106).
// Arithmetic is fp64 like the reference (the 6x6 / 7x6 solves are too ill-conditioned for fp32 at the stated
// 1e-5 tolerance); I/O is fp32.  Screw axes w0/p0 and the site pose at q = 0 come from the compiled model.
#pragma once
#include "avsim_dev.h"

struct DiffIKParams {
    float k_pos, k_ori, damping, max_angvel, dt, k_null[7], q0[7];
    int iterations;
};
struct GradIKParams {
    float step_size, min_cost_delta, position_weight, rotation_weight, position_threshold, rotation_threshold,
        max_pos_diff, max_rot_diff, joint_p, center_w[7], disp_w[7];
    int max_iterations;
};

struct T44 {
    double R[9], p[3];
};
__device__ inline void t_mul(const T44 &a, const T44 &b, T44 &o) {
    T44 r;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) r.R[3 * i + j] = a.R[3 * i] * b.R[j] + a.R[3 * i + 1] * b.R[3 + j] + a.R[3 * i + 2] * b.R[6 + j];
        r.p[i] = a.R[3 * i] * b.p[0] + a.R[3 * i + 1] * b.p[1] + a.R[3 * i + 2] * b.p[2] + a.p[i];
    }
    o = r;
}
// exp2mat (transform_utils.py:239-261) for a unit rotation axis w, v = -w x p
__device__ inline void exp2mat(const double *w, const double *v, double th, T44 &T) {
    double s = sin(th), c = cos(th);
    double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0}, W2[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) W2[3 * i + j] = W[3 * i] * W[j] + W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j];
    for (int i = 0; i < 9; i++) T.R[i] = ((i % 4) == 0 ? 1.0 : 0.0) + s * W[i] + (1 - c) * W2[i];
    for (int i = 0; i < 3; i++) {
        double acc = 0;
        for (int j = 0; j < 3; j++) acc += (((i == j) ? th : 0.0) + (1 - c) * W[3 * i + j] + (th - s) * W2[3 * i + j]) * v[j];
        T.p[i] = acc;
    }
}
struct ArmTab {
    int n;
    double w[7][3], v[7][3], lo[7], hi[7];
    T44 site0;
};
__device__ inline void load_arm(const DevModel &m, int arm, ArmTab &A) {
    A.n = m.ik_ndof[arm];
    for (int i = 0; i < 7; i++) {
        const float *w = m.ik_w0 + (arm * 7 + i) * 3, *p = m.ik_p0 + (arm * 7 + i) * 3;
        for (int k = 0; k < 3; k++) A.w[i][k] = w[k];
        A.v[i][0] = -(w[1] * (double)p[2] - w[2] * (double)p[1]);
        A.v[i][1] = -(w[2] * (double)p[0] - w[0] * (double)p[2]);
        A.v[i][2] = -(w[0] * (double)p[1] - w[1] * (double)p[0]);
        A.lo[i] = m.ik_range[(arm * 7 + i) * 2]; A.hi[i] = m.ik_range[(arm * 7 + i) * 2 + 1];
    }
    const float *s = m.ik_site0 + arm * 16;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) A.site0.R[3 * i + j] = s[4 * i + j]; A.site0.p[i] = s[4 * i + 3]; }
}
// forward_kinematics (kinematics.py:17-24)
__device__ inline void ik_fk(const ArmTab &A, const double *th, T44 &M) {
    M = A.site0;
    for (int i = A.n - 1; i >= 0; i--) {
        T44 T;
        exp2mat(A.w[i], A.v[i], th[i], T);
        t_mul(T, M, M);
    }
}
// jacobian (kinematics.py:35-50): columns Ad(T_{i-1}) S_i, rows reordered to [v; w]
__device__ inline void ik_jac(const ArmTab &A, const double *th, double J[6][7]) {
    T44 Ts;
    for (int i = 0; i < 9; i++) Ts.R[i] = (i % 4) == 0;
    Ts.p[0] = Ts.p[1] = Ts.p[2] = 0;
    for (int i = 0; i < A.n; i++) {
        double Rw[3], Rv[3];
        for (int r = 0; r < 3; r++) {
            Rw[r] = Ts.R[3 * r] * A.w[i][0] + Ts.R[3 * r + 1] * A.w[i][1] + Ts.R[3 * r + 2] * A.w[i][2];
            Rv[r] = Ts.R[3 * r] * A.v[i][0] + Ts.R[3 * r + 1] * A.v[i][1] + Ts.R[3 * r + 2] * A.v[i][2];
        }
        J[3][i] = Rw[0]; J[4][i] = Rw[1]; J[5][i] = Rw[2];
        J[0][i] = Ts.p[1] * Rw[2] - Ts.p[2] * Rw[1] + Rv[0];
        J[1][i] = Ts.p[2] * Rw[0] - Ts.p[0] * Rw[2] + Rv[1];
        J[2][i] = Ts.p[0] * Rw[1] - Ts.p[1] * Rw[0] + Rv[2];
        T44 E;
        exp2mat(A.w[i], A.v[i], th[i], E);
        t_mul(Ts, E, Ts);
    }
}
// quat2mat (transform_utils.py:52-79) on an (x,y,z,w) quaternion, with the reference's float32 round trip
__device__ inline void ik_quat2mat_xyzw(const double *qxyzw, double *R) {
    float q[4] = {(float)qxyzw[3], (float)qxyzw[0], (float)qxyzw[1], (float)qxyzw[2]};
    float n = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (n < 8.881784197001252e-16f) { for (int i = 0; i < 9; i++) R[i] = (i % 4) == 0; return; }
    double s = sqrt(2.0 / (double)n);
    for (int i = 0; i < 4; i++) q[i] = (float)((double)q[i] * s);
    float q2[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) q2[i][j] = q[i] * q[j];
    // the reference's off-diagonal sums are float32 +/- float32 (numpy promotion keeps float32); the diagonal starts
    // from the float64 literal 1.0 and is therefore float64
    R[0] = 1.0 - q2[2][2] - q2[3][3]; R[1] = (double)(q2[1][2] - q2[3][0]); R[2] = (double)(q2[1][3] + q2[2][0]);
    R[3] = (double)(q2[1][2] + q2[3][0]); R[4] = 1.0 - q2[1][1] - q2[3][3]; R[5] = (double)(q2[2][3] - q2[1][0]);
    R[6] = (double)(q2[1][3] - q2[2][0]); R[7] = (double)(q2[2][3] + q2[1][0]); R[8] = 1.0 - q2[1][1] - q2[2][2];
}
// angular_error (transform_utils.py:183-194)
__device__ inline void ik_ang_err(const double *des, const double *cur, double *e) {
    e[0] = e[1] = e[2] = 0;
    for (int c = 0; c < 3; c++) {
        double rc[3] = {cur[c], cur[3 + c], cur[6 + c]}, rd[3] = {des[c], des[3 + c], des[6 + c]};
        e[0] += 0.5 * (rc[1] * rd[2] - rc[2] * rd[1]);
        e[1] += 0.5 * (rc[2] * rd[0] - rc[0] * rd[2]);
        e[2] += 0.5 * (rc[0] * rd[1] - rc[1] * rd[0]);
    }
}
// solve the SPD system A x = b (n <= 7) by Cholesky; A is overwritten
__device__ inline void spd_solve(double A[7][7], int n, double *b) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i][j];
            for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k];
            A[i][j] = (i == j) ? sqrt(s) : s / A[j][j];
        }
    for (int i = 0; i < n; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= A[i][k] * b[k]; b[i] = s / A[i][i]; }
    for (int i = n - 1; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < n; k++) s -= A[k][i] * b[k]; b[i] = s / A[i][i]; }
}

// x = pinv(G) b for a symmetric positive SEMI-definite 6x6 G (= J J'), the way np.linalg.pinv treats a rank-deficient
// Jacobian (reference diff_ik.py:72: directions whose singular value is below the cutoff are dropped, not divided by):
// cyclic Jacobi eigen-decomposition, eigenvalues below 1e-14 of the largest -- the resolution of the squared spectrum in fp64 --
// count as zero.  Only taken when the Cholesky path meets a vanishing pivot (at / next to a kinematic singularity).
__device__ inline void psd6_pinv_apply(double G[7][7], double *b) {
    double V[6][6];
    for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; sweep++) {
        double off = 0;
        for (int p = 0; p < 6; p++) for (int q = p + 1; q < 6; q++) off += G[p][q] * G[p][q];
        if (off < 1e-40) break;
        for (int p = 0; p < 6; p++)
            for (int q = p + 1; q < 6; q++) {
                if (fabs(G[p][q]) < 1e-300) continue;
                double th = (G[q][q] - G[p][p]) / (2.0 * G[p][q]);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0)), c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 6; k++) { double a = G[k][p], bb = G[k][q]; G[k][p] = c * a - sn * bb; G[k][q] = sn * a + c * bb; }
                for (int k = 0; k < 6; k++) { double a = G[p][k], bb = G[q][k]; G[p][k] = c * a - sn * bb; G[q][k] = sn * a + c * bb; }
                for (int k = 0; k < 6; k++) { double a = V[k][p], bb = V[k][q]; V[k][p] = c * a - sn * bb; V[k][q] = sn * a + c * bb; }
            }
    }
    double lmax = 0;
    for (int i = 0; i < 6; i++) lmax = fmax(lmax, G[i][i]);
    double x[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 6; i++) {
        if (!(G[i][i] > 1e-14 * lmax)) continue;
        double s = 0;
        for (int k = 0; k < 6; k++) s += V[k][i] * b[k];
        s /= G[i][i];
        for (int k = 0; k < 6; k++) x[k] += V[k][i] * s;
    }
    for (int k = 0; k < 6; k++) b[k] = x[k];
}
// Cholesky solve of the 6x6 system; false (A, b untouched by the caller's copy) when a pivot vanishes
__device__ inline bool spd6_solve_checked(double A[7][7], double *b) {
    double dmax = 0;
    for (int i = 0; i < 6; i++) dmax = fmax(dmax, A[i][i]);
    for (int i = 0; i < 6; i++)
        for (int j = 0; j <= i; j++) {
            double s = A[i][j];
            for (int k = 0; k < j; k++) s -= A[i][k] * A[j][k];
            if (i == j) {
                if (!(s > 1e-12 * dmax)) return false;
                A[i][j] = sqrt(s);
            } else
                A[i][j] = s / A[j][j];
        }
    for (int i = 0; i < 6; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= A[i][k] * b[k]; b[i] = s / A[i][i]; }
    for (int i = 5; i >= 0; i--) { double s = b[i]; for (int k = i + 1; k < 6; k++) s -= A[k][i] * b[k]; b[i] = s / A[i][i]; }
    return true;
}

__global__ void avsim_fk_kernel(DevModel m, int arm, const float *__restrict__ q, int n, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ArmTab A;
    load_arm(m, arm, A);
    double th[7];
    for (int k = 0; k < A.n; k++) th[k] = q[(size_t)i * A.n + k];
    T44 T;
    ik_fk(A, th, T);
    float *o = out + (size_t)i * 16;
    for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) o[4 * r + c] = (float)T.R[3 * r + c]; o[4 * r + 3] = (float)T.p[r]; }
    o[12] = o[13] = o[14] = 0.f; o[15] = 1.f;
}

// jacobian(theta) of create_jac_fn (kinematics.py:28-52): space Jacobian, rows [v; w]; out f32 [n][6][ndof]
__global__ void avsim_jac_kernel(DevModel m, int arm, const float *__restrict__ q, int n, float *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ArmTab A;
    load_arm(m, arm, A);
    double th[7], J[6][7];
    for (int k = 0; k < A.n; k++) th[k] = q[(size_t)i * A.n + k];
    ik_jac(A, th, J);
    float *o = out + (size_t)i * 6 * A.n;
    for (int r = 0; r < 6; r++)
        for (int c = 0; c < A.n; c++) o[r * A.n + c] = (float)J[r][c];
}

// DiffIK.run (diff_ik.py:51-90)
__global__ void avsim_diffik_kernel(DevModel m, int arm, const float *__restrict__ q_in, const float *__restrict__ pos,
                                    const float *__restrict__ quat_wxyz, int n, DiffIKParams P, float *__restrict__ q_out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    ArmTab A;
    load_arm(m, arm, A);
    int nd = A.n;
    double q[7], tp[3], tq[4], Rt[9];
    for (int k = 0; k < nd; k++) q[k] = q_in[(size_t)idx * nd + k];
    for (int k = 0; k < 3; k++) tp[k] = pos[(size_t)idx * 3 + k];
    tq[0] = quat_wxyz[(size_t)idx * 4 + 1]; tq[1] = quat_wxyz[(size_t)idx * 4 + 2]; tq[2] = quat_wxyz[(size_t)idx * 4 + 3];
    tq[3] = quat_wxyz[(size_t)idx * 4];
    ik_quat2mat_xyzw(tq, Rt);
    double dt = P.dt;
    for (int it = 0; it < P.iterations; it++) {
        T44 T;
        ik_fk(A, q, T);
        double twist[6], dr[3];
        for (int k = 0; k < 3; k++) twist[k] = (double)P.k_pos * (tp[k] - T.p[k]) / dt;
        ik_ang_err(Rt, T.R, dr);
        for (int k = 0; k < 3; k++) twist[3 + k] = (double)P.k_ori * dr[k] / dt;
        double J[6][7];
        ik_jac(A, q, J);
        // dq = J^T (J J^T + damping I)^-1 twist
        double G[7][7], y[7], dq[7];
        for (int i = 0; i < 6; i++) {
            for (int j = 0; j < 6; j++) {
                double s = 0;
                for (int k = 0; k < nd; k++) s += J[i][k] * J[j][k];
                G[i][j] = s + (i == j ? (double)P.damping : 0.0);
            }
            y[i] = twist[i];
        }
        spd_solve(G, 6, y);
        for (int k = 0; k < nd; k++) { double s = 0; for (int i = 0; i < 6; i++) s += J[i][k] * y[i]; dq[k] = s; }
        // null space: (I - pinv(J) J) (k_null * (q0 - q)),  pinv(J) = J^T (J J^T)^-1 for full row rank
        double z[7], Jz[7];
        for (int k = 0; k < nd; k++) z[k] = (double)P.k_null[k] * ((double)P.q0[k] - q[k]);
        for (int i = 0; i < 6; i++) {
            double s = 0;
            for (int k = 0; k < nd; k++) s += J[i][k] * z[k];
            Jz[i] = s;
            for (int j = 0; j < 6; j++) { double g = 0; for (int k = 0; k < nd; k++) g += J[i][k] * J[j][k]; G[i][j] = g; }
        }
        {
            double G2[7][7], Jz2[7];
            for (int i = 0; i < 6; i++) { Jz2[i] = Jz[i]; for (int j = 0; j < 6; j++) G2[i][j] = G[i][j]; }
            if (spd6_solve_checked(G2, Jz2)) { for (int i = 0; i < 6; i++) Jz[i] = Jz2[i]; }
            else psd6_pinv_apply(G, Jz);     // rank-deficient Jacobian: thresholded pseudo-inverse, finite like np.linalg.pinv
        }
        for (int k = 0; k < nd; k++) {
            double s = 0;
            for (int i = 0; i < 6; i++) s += J[i][k] * Jz[i];
            dq[k] += z[k] - s;
        }
        for (int k = 0; k < nd; k++) {
            double d = fmin(fmax(dq[k], -(double)P.max_angvel), (double)P.max_angvel);
            q[k] = fmin(fmax(q[k] + d * dt, A.lo[k]), A.hi[k]);
        }
    }
    for (int k = 0; k < nd; k++) q_out[(size_t)idx * nd + k] = (float)q[k];
}

// ---- GradIK helpers
// rotation matrix -> (x,y,z,w) quaternion with w >= 0 (mat2quat, transform_utils.py:9-49, eigenvector form)
__device__ inline void ik_mat2quat_xyzw(const double *R, double *q) {
    double tr = R[0] + R[4] + R[8], w, x, y, z;
    if (tr > 0) { double s = sqrt(tr + 1.0) * 2; w = 0.25 * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s; }
    else if (R[0] > R[4] && R[0] > R[8]) { double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; w = (R[7] - R[5]) / s; x = 0.25 * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s; }
    else if (R[4] > R[8]) { double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25 * s; z = (R[5] + R[7]) / s; }
    else { double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25 * s; }
    if (w < 0) { w = -w; x = -x; y = -y; z = -z; }
    double nn = sqrt(w * w + x * x + y * y + z * z);
    q[0] = x / nn; q[1] = y / nn; q[2] = z / nn; q[3] = w / nn;
}
// quat2axisangle / axisangle2quat (transform_utils.py:82-133) on (x,y,z,w); np.isclose(x, 0) = |x| <= 1e-8
__device__ inline void ik_quat2axisangle(const double *q, double *rv) {
    double wq = fmin(fmax(q[3], -1.0), 1.0), den = sqrt(1.0 - wq * wq);
    rv[0] = rv[1] = rv[2] = 0;
    if (den > 1e-8) { double sc = 2.0 * acos(wq) / den; rv[0] = q[0] * sc; rv[1] = q[1] * sc; rv[2] = q[2] * sc; }
}
__device__ inline void ik_axisangle2quat(const double *v, double *q) {
    double a = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    q[0] = q[1] = q[2] = 0; q[3] = 1;
    if (a > 1e-8) { double s = sin(a / 2) / a; q[0] = v[0] * s; q[1] = v[1] * s; q[2] = v[2] * s; q[3] = cos(a / 2); }
}
// limit_pose (transform_utils.py:263-287): the target (tp, tR) is pulled to within max_pos / max_rot of the current pose, in place
__device__ inline void ik_limit_pose(const double *cp, const double *cR, double *tp, double *tR, double max_pos, double max_rot) {
    double d[3] = {tp[0] - cp[0], tp[1] - cp[1], tp[2] - cp[2]};
    double dn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (dn > max_pos) for (int k = 0; k < 3; k++) d[k] = d[k] / dn * max_pos;
    for (int k = 0; k < 3; k++) tp[k] = cp[k] + d[k];
    double Rrel[9];
    for (int i = 0; i < 3; i++)   // target @ inv(current) = target @ current^T
        for (int j = 0; j < 3; j++) Rrel[3 * i + j] = tR[3 * i] * cR[3 * j] + tR[3 * i + 1] * cR[3 * j + 1] + tR[3 * i + 2] * cR[3 * j + 2];
    double qr[4], rv[3];
    ik_mat2quat_xyzw(Rrel, qr);
    ik_quat2axisangle(qr, rv);
    double ang = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    if (ang > max_rot) {
        double sc = max_rot / ang, v[3] = {rv[0] * sc, rv[1] * sc, rv[2] * sc}, ql[4];
        ik_axisangle2quat(v, ql);
        double Rl[9], Rn[9];
        ik_quat2mat_xyzw(ql, Rl);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) Rn[3 * i + j] = Rl[3 * i] * cR[j] + Rl[3 * i + 1] * cR[3 + j] + Rl[3 * i + 2] * cR[6 + j];
        for (int i = 0; i < 9; i++) tR[i] = Rn[i];
    }
}
__device__ inline double gradik_cost(const ArmTab &A, const GradIKParams &P, const double *q, const double *q_start,
                                     const double *tp, const double *tR, const double *cw, const double *centers) {
    T44 T;
    ik_fk(A, q, T);
    double d[3] = {tp[0] - T.p[0], tp[1] - T.p[1], tp[2] - T.p[2]}, e[3];
    ik_ang_err(tR, T.R, e);
    double pw = (double)P.position_weight * sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double rw = (double)P.rotation_weight * sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    double cost = pw * pw + rw * rw;
    for (int k = 0; k < A.n; k++) {
        double c = cw[k] * (q[k] - centers[k]), dd = (double)P.disp_w[k] * (q[k] - q_start[k]);
        cost += c * c + dd * dd;
    }
    return cost;
}

// GradIK.run (grad_ik.py:8-99,150-166)
__global__ void avsim_gradik_kernel(DevModel m, int arm, const float *__restrict__ q_in, const float *__restrict__ pos,
                                    const float *__restrict__ quat_wxyz, int n, GradIKParams P, float *__restrict__ q_out) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    ArmTab A;
    load_arm(m, arm, A);
    int nd = A.n;
    double q0[7], tp[3], tq[4], tR[9], cw[7], centers[7];
    for (int k = 0; k < nd; k++) {
        q0[k] = q_in[(size_t)idx * nd + k];
        centers[k] = 0.5 * (A.lo[k] + A.hi[k]);
        cw[k] = (double)P.center_w[k];   // already divided by the half range on the host (grad_ik.py:137)
    }
    for (int k = 0; k < 3; k++) tp[k] = pos[(size_t)idx * 3 + k];
    tq[0] = quat_wxyz[(size_t)idx * 4 + 1]; tq[1] = quat_wxyz[(size_t)idx * 4 + 2]; tq[2] = quat_wxyz[(size_t)idx * 4 + 3];
    tq[3] = quat_wxyz[(size_t)idx * 4];
    ik_quat2mat_xyzw(tq, tR);
    // limit_pose (transform_utils.py:263-287)
    T44 T;
    ik_fk(A, q0, T);
    ik_limit_pose(T.p, T.R, tp, tR, (double)P.max_pos_diff, (double)P.max_rot_diff);
    double step = (double)P.step_size;
    double init_cost = gradik_cost(A, P, q0, q0, tp, tR, cw, centers);
    double grad[7], working[7], local[7], best[7];
    for (int k = 0; k < nd; k++) working[k] = local[k] = best[k] = q0[k];
    double local_cost = init_cost, best_cost = init_cost, previous_cost = 0.0;
    for (int it = 0; it < P.max_iterations; it++) {
        for (int i = 0; i < nd; i++) {
            working[i] = local[i] - step;
            double p1 = gradik_cost(A, P, working, q0, tp, tR, cw, centers);
            working[i] = local[i] + step;
            double p3 = gradik_cost(A, P, working, q0, tp, tR, cw, centers);
            working[i] = local[i];
            grad[i] = p3 - p1;
        }
        double sum = step;
        for (int i = 0; i < nd; i++) sum += fabs(grad[i]);
        double f = step / sum;
        for (int i = 0; i < nd; i++) grad[i] *= f;
        for (int i = 0; i < nd; i++) working[i] = local[i] - grad[i];
        double p1 = gradik_cost(A, P, working, q0, tp, tR, cw, centers);
        for (int i = 0; i < nd; i++) working[i] = local[i] + grad[i];
        double p3 = gradik_cost(A, P, working, q0, tp, tR, cw, centers);
        double p2 = 0.5 * (p1 + p3), cost_diff = 0.5 * (p3 - p1);
        double joint_diff = (isfinite(cost_diff) && cost_diff != 0.0) ? p2 / cost_diff : 0.0;
        for (int i = 0; i < nd; i++) {
            working[i] = fmin(fmax(local[i] - grad[i] * joint_diff, A.lo[i]), A.hi[i]);
            local[i] = working[i];
        }
        local_cost = gradik_cost(A, P, local, q0, tp, tR, cw, centers);
        if (local_cost < best_cost) { for (int i = 0; i < nd; i++) best[i] = local[i]; best_cost = local_cost; }
        // solution_fn (grad_ik.py:205-218): within_pose_threshold
        ik_fk(A, local, T);
        double d[3] = {tp[0] - T.p[0], tp[1] - T.p[1], tp[2] - T.p[2]}, e[3];
        ik_ang_err(tR, T.R, e);
        if (sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < (double)P.position_threshold &&
            sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) < (double)P.rotation_threshold)
            break;
        if (fabs(local_cost - previous_cost) <= (double)P.min_cost_delta) break;
        previous_cost = local_cost;
    }
    for (int k = 0; k < nd; k++) q_out[(size_t)idx * nd + k] = (float)(q0[k] + (double)P.joint_p * (best[k] - q0[k]));
}

// GradIK.run, low-latency form: 16 lanes per problem.  One iteration of the descent is 2 ndof + 2 + 1 cost evaluations and one
// FK that a single thread walks one after the other (16 forward-kinematics chains per iteration, 50 iterations); here lane
// 2i / 2i+1 of the group evaluate the -/+ finite difference of joint i at the same time, lanes 0 / 1 the two probes, lane 0 the
// cost of the new iterate while lane 1 does the pose-threshold FK: 3 chains per iteration instead of 16.  Every evaluation is done
// entirely by one lane with the arithmetic of the thread-per-problem kernel and only finished doubles are exchanged
// (__shfl_sync, width 16), so the two kernels return the same iterates.  Two problems share a warp and may leave the loop at
// different iterations: all shuffles are masked to the problem's own half-warp.
#if defined(__CUDACC__)   // device build only: the host emulation harness (tests/emu) runs the thread-per-problem form
__global__ void avsim_gradik_group_kernel(DevModel m, int arm, const float *__restrict__ q_in, const float *__restrict__ pos,
                                          const float *__restrict__ quat_wxyz, int n, GradIKParams P, float *__restrict__ q_out) {
    const int gidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, sub = threadIdx.x & 15;
    const unsigned gm = 0xffffu << (threadIdx.x & 16);
    const bool valid = gidx < n;
    const int idx = valid ? gidx : n - 1;   // surplus groups recompute the last problem and do not store
    ArmTab A;
    load_arm(m, arm, A);
    int nd = A.n;
    double q0[7], tp[3], tq[4], tR[9], cw[7], centers[7];
    for (int k = 0; k < nd; k++) {
        q0[k] = q_in[(size_t)idx * nd + k];
        centers[k] = 0.5 * (A.lo[k] + A.hi[k]);
        cw[k] = (double)P.center_w[k];
    }
    for (int k = 0; k < 3; k++) tp[k] = pos[(size_t)idx * 3 + k];
    tq[0] = quat_wxyz[(size_t)idx * 4 + 1]; tq[1] = quat_wxyz[(size_t)idx * 4 + 2]; tq[2] = quat_wxyz[(size_t)idx * 4 + 3];
    tq[3] = quat_wxyz[(size_t)idx * 4];
    ik_quat2mat_xyzw(tq, tR);
    T44 T;
    ik_fk(A, q0, T);
    ik_limit_pose(T.p, T.R, tp, tR, (double)P.max_pos_diff, (double)P.max_rot_diff);
    double step = (double)P.step_size;
    double init_cost = gradik_cost(A, P, q0, q0, tp, tR, cw, centers);
    double grad[7], working[7], local[7], best[7];
    for (int k = 0; k < nd; k++) working[k] = local[k] = best[k] = q0[k];
    double local_cost = init_cost, best_cost = init_cost, previous_cost = 0.0;
    for (int it = 0; it < P.max_iterations; it++) {
        double mine = 0.0;
        if (sub < 2 * nd) {
            const int i = sub >> 1;
            working[i] = (sub & 1) ? local[i] + step : local[i] - step;
            mine = gradik_cost(A, P, working, q0, tp, tR, cw, centers);
            working[i] = local[i];
        }
        for (int i = 0; i < nd; i++) {
            double p1 = __shfl_sync(gm, mine, 2 * i, 16), p3 = __shfl_sync(gm, mine, 2 * i + 1, 16);
            grad[i] = p3 - p1;
        }
        double sum = step;
        for (int i = 0; i < nd; i++) sum += fabs(grad[i]);
        double f = step / sum;
        for (int i = 0; i < nd; i++) grad[i] *= f;
        if (sub < 2) {
            for (int i = 0; i < nd; i++) working[i] = sub ? local[i] + grad[i] : local[i] - grad[i];
            mine = gradik_cost(A, P, working, q0, tp, tR, cw, centers);
        }
        double p1 = __shfl_sync(gm, mine, 0, 16), p3 = __shfl_sync(gm, mine, 1, 16);
        double p2 = 0.5 * (p1 + p3), cost_diff = 0.5 * (p3 - p1);
        double joint_diff = (isfinite(cost_diff) && cost_diff != 0.0) ? p2 / cost_diff : 0.0;
        for (int i = 0; i < nd; i++) {
            working[i] = fmin(fmax(local[i] - grad[i] * joint_diff, A.lo[i]), A.hi[i]);
            local[i] = working[i];
        }
        int within = 0;
        if (sub == 0) mine = gradik_cost(A, P, local, q0, tp, tR, cw, centers);
        if (sub == 1) {   // solution_fn (grad_ik.py:205-218): within_pose_threshold
            ik_fk(A, local, T);
            double d[3] = {tp[0] - T.p[0], tp[1] - T.p[1], tp[2] - T.p[2]}, e[3];
            ik_ang_err(tR, T.R, e);
            within = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < (double)P.position_threshold &&
                     sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) < (double)P.rotation_threshold;
        }
        local_cost = __shfl_sync(gm, mine, 0, 16);
        within = __shfl_sync(gm, within, 1, 16);
        if (local_cost < best_cost) { for (int i = 0; i < nd; i++) best[i] = local[i]; best_cost = local_cost; }
        if (within) break;
        if (fabs(local_cost - previous_cost) <= (double)P.min_cost_delta) break;
        previous_cost = local_cost;
    }
    if (valid && sub < nd) q_out[(size_t)idx * nd + sub] = (float)(q0[sub] + (double)P.joint_p * (best[sub] - q0[sub]));
}
#endif

// ---- transform_utils.py primitives as a batched operator (fp64 in / out, one thread per item): the callable mirror of the
// helpers the two controllers use internally, so that each can be checked on its own against the reference's numba functions
// (tests/test_transform_utils.py) and used by host code on whole batches of poses (av_aloha_b200/transform_utils.py).
enum { XF_MAT2QUAT = 0, XF_QUAT2MAT, XF_QUAT2AXISANGLE, XF_AXISANGLE2QUAT, XF_ANGULAR_ERROR, XF_LIMIT_POSE, XF_EXP2MAT, XF_ADJOINT,
       XF_WITHIN_POSE, XF_NOPS };
__global__ void avsim_transform_kernel(int op, const double *__restrict__ a, const double *__restrict__ b, int n, double p0, double p1,
                                       double *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    switch (op) {
    case XF_MAT2QUAT: ik_mat2quat_xyzw(a + 9 * (size_t)i, out + 4 * (size_t)i); break;
    case XF_QUAT2MAT: ik_quat2mat_xyzw(a + 4 * (size_t)i, out + 9 * (size_t)i); break;
    case XF_QUAT2AXISANGLE: ik_quat2axisangle(a + 4 * (size_t)i, out + 3 * (size_t)i); break;
    case XF_AXISANGLE2QUAT: ik_axisangle2quat(a + 3 * (size_t)i, out + 4 * (size_t)i); break;
    case XF_ANGULAR_ERROR: ik_ang_err(a + 9 * (size_t)i, b + 9 * (size_t)i, out + 3 * (size_t)i); break;
    case XF_LIMIT_POSE: {   // a = current [pos 3 | mat 9], b = target [pos 3 | mat 9], p0 = max_pos_diff, p1 = max_rot_diff
        double tp[3], tR[9];
        for (int k = 0; k < 3; k++) tp[k] = b[12 * (size_t)i + k];
        for (int k = 0; k < 9; k++) tR[k] = b[12 * (size_t)i + 3 + k];
        ik_limit_pose(a + 12 * (size_t)i, a + 12 * (size_t)i + 3, tp, tR, p0, p1);
        for (int k = 0; k < 3; k++) out[12 * (size_t)i + k] = tp[k];
        for (int k = 0; k < 9; k++) out[12 * (size_t)i + 3 + k] = tR[k];
        break;
    }
    case XF_EXP2MAT: {      // a = [w 3 | v 3 | theta]; |w| = 1, or w = 0 and |v| = 1 (pure translation)
        const double *x = a + 7 * (size_t)i;
        T44 T;
        if (x[0] * x[0] + x[1] * x[1] + x[2] * x[2] <= 1e-16) {
            for (int k = 0; k < 9; k++) T.R[k] = (k % 4) == 0;
            for (int k = 0; k < 3; k++) T.p[k] = x[3 + k] * x[6];
        } else
            exp2mat(x, x + 3, x[6], T);
        double *o = out + 16 * (size_t)i;
        for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) o[4 * r + c] = T.R[3 * r + c]; o[4 * r + 3] = T.p[r]; }
        o[12] = o[13] = o[14] = 0; o[15] = 1;
        break;
    }
    case XF_ADJOINT: {      // a = 4x4 T -> 6x6 [[R 0] [skew(p) R, R]]
        const double *T = a + 16 * (size_t)i;
        double *o = out + 36 * (size_t)i;
        double R[9], p[3] = {T[3], T[7], T[11]};
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[3 * r + c] = T[4 * r + c];
        double S[9] = {0, -p[2], p[1], p[2], 0, -p[0], -p[1], p[0], 0};
        for (int k = 0; k < 36; k++) o[k] = 0;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) {
                o[6 * r + c] = R[3 * r + c];
                o[6 * (r + 3) + c + 3] = R[3 * r + c];
                o[6 * (r + 3) + c] = S[3 * r] * R[c] + S[3 * r + 1] * R[3 + c] + S[3 * r + 2] * R[6 + c];
            }
        break;
    }
    case XF_WITHIN_POSE: {  // a = current [pos | mat], b = target [pos | mat], p0 / p1 = position / rotation threshold
        const double *c = a + 12 * (size_t)i, *t = b + 12 * (size_t)i;
        double d[3] = {t[0] - c[0], t[1] - c[1], t[2] - c[2]}, e[3];
        ik_ang_err(t + 3, c + 3, e);
        out[i] = (sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) < p0 && sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) < p1) ? 1.0 : 0.0;
        break;
    }
    default: break;
    }
}
