/*
 * hk_oracle_lqng.c — CPU ORACLE (test infrastructure, NOT the product).
 *
 * Plain-C restatement of the reference's feedback LQ Nash game solver, dense matrices, double precision, in the
 * reference's own association order, quirks included:
 *   KartLQR.solveFeedbackLQR                  Assets/Karting/Scripts/AI/LQR/KartLQR.cs:17-128
 *   LinearizedBicycle.getA / getB             Assets/Karting/Scripts/AI/LQR/KartLQRDynamics.cs:40-62
 *   LQRCheckpointReachAvoidCost.getQMatrix /
 *     getQVec / getRMatrix                    Assets/Karting/Scripts/AI/LQR/KartLQRCosts.cs:57-140
 * The arithmetic of the reference lives in MathNet.Numerics 4.15.0 (binary only, Assets/Plugins/MathNet.Numerics.dll):
 * Matrix.Solve on a square matrix = LU with partial pivoting (JAMA-style Doolittle, restated in lu_factor below).
 *
 * PARITY UNPINNED: the reference ships no tests/golden vectors and cannot be compiled here (no .NET in this image or
 * on the GPU box — probed, gpurun_out/mb1.log: "no-dotnet").  What pins this file is (a) an independent numpy
 * restatement (oracle/np_lqng.py, LAPACK solve) and (b) the surveyor's smoke value.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include "hk_oracle.h"

#define XD 4
#define UD 2
#define NMAX 16
#define MMAX 8

/* KartLQRDynamics.cs:40-51 — A_i = I + dt * d f/dx about x0 = (x, z, v, h) */
void hk_oracle_bicycle_A(double dt, const double* x0, double* A /*[4][4]*/)
{
    memset(A, 0, sizeof(double) * XD * XD);
    for (int i = 0; i < XD; ++i) A[i * XD + i] = 1.0;             /* SparseIdentity :44 */
    A[0 * XD + 2] = cos(x0[3]) * dt;                               /* [x, v] :45 */
    A[1 * XD + 2] = sin(x0[3]) * dt;                               /* [z, v] :46 */
    A[0 * XD + 3] = -sin(x0[3]) * dt * x0[2];                      /* [x, h] :47 */
    A[1 * XD + 3] = cos(x0[3]) * dt * x0[2];                       /* [z, h] :48 */
}

/* KartLQRDynamics.cs:53-62 */
void hk_oracle_bicycle_B(double dt, double* B /*[4][2]*/)
{
    memset(B, 0, sizeof(double) * XD * UD);
    B[2 * UD + 0] = dt;                                            /* [v, 0] :58 */
    B[3 * UD + 1] = dt;                                            /* [h, 1] :59 */
}

/*
 * KartLQRCosts.cs:57-98 (getQMatrix), :103-127 (getQVec), :132-140 (getRMatrix) for one player whose
 * avoidDynamics list has n_other entries (private ordering [self, others...]; every block is 4 wide).
 *   target[4], tw[4] (targetWeights for x,z,v,h — all four keys are present in HierarchicalKartAgent.cs:930-962)
 *   aw[n_other][2]   avoidWeights[x][k], avoidWeights[z][k]   (avoidIndices are x->x, z->z: HKA:1013-1023)
 *   otgt[n_other][4] opponentTargetStates, otw[n_other][3] opponentTargetWeights for x,z,v (HKA:1071-1091)
 */
void hk_oracle_cost(int n_other, const double* target, const double* tw, double cw,
                    const double* aw, const double* otgt, const double* otw,
                    double* Q /*[n][n]*/, double* q /*[n]*/, double* R /*[2][2]*/)
{
    const int n = XD * (1 + n_other);
    memset(Q, 0, sizeof(double) * n * n);
    /* avoid terms: foreach currStateIndex in avoidWeights.Keys (x then z) :64-80 */
    for (int s = 0; s < 2; ++s) {
        int curr = XD;
        double total = 0.0;
        for (int k = 0; k < n_other; ++k) {
            int t = curr + s;                                      /* idxIndices[k] == s */
            double w = aw[k * 2 + s];
            Q[s * n + t] = w;                                      /* :73 */
            Q[t * n + s] = w;                                      /* :74 */
            Q[t * n + t] = -w;                                     /* :75 */
            total -= w;                                            /* :76 */
            curr += XD;
        }
        Q[s * n + s] = total;                                      /* :79 */
    }
    for (int s = 0; s < XD; ++s) Q[s * n + s] += tw[s];            /* own target weights are ADDED :81-84 */
    {
        int curr = XD;
        for (int k = 0; k < n_other; ++k) {
            for (int o = 0; o < 3; ++o)                            /* keys x, z, v */
                Q[(curr + o) * n + (curr + o)] = -otw[k * 3 + o];  /* ASSIGNED, overwriting the avoid diagonal :91 (quirk Q4) */
            curr += XD;
        }
    }
    /* getQVec :103-127 */
    memset(q, 0, sizeof(double) * n);
    for (int s = 0; s < XD; ++s) q[s] = -target[s];                /* :109 */
    for (int s = 0; s < XD; ++s) q[s] = q[s] * tw[s];              /* :110-113 */
    {
        int curr = XD;
        for (int k = 0; k < n_other; ++k) {
            for (int s = 0; s < XD; ++s) q[curr + s] = otgt[k * XD + s];          /* :117 */
            for (int o = 0; o < 3; ++o) q[curr + o] = q[curr + o] * -otw[k * 3 + o]; /* :121 */
            curr += XD;
        }
    }
    R[0] = cw; R[1] = 0.0; R[2] = 0.0; R[3] = cw;                  /* SparseIdentity * controlWeight :136 */
}

/* ---- small dense helpers (row-major) ------------------------------------------------------------------------ */
static void matmul(int r, int k, int c, const double* X, const double* Y, double* Z)
{
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += X[i * k + l] * Y[l * c + j];
            Z[i * c + j] = s;
        }
}
/* Z = X^T Y, X is k x r, Y is k x c  (MathNet TransposeThisAndMultiply) */
static void matmul_tn(int k, int r, int c, const double* X, const double* Y, double* Z)
{
    for (int i = 0; i < r; ++i)
        for (int j = 0; j < c; ++j) {
            double s = 0.0;
            for (int l = 0; l < k; ++l) s += X[l * r + i] * Y[l * c + j];
            Z[i * c + j] = s;
        }
}

/* LU with partial pivoting, JAMA / MathNet managed "LUFactor" column algorithm. Returns 1 if a pivot is exactly 0. */
static int lu_factor(int m, double* LU, int* piv)
{
    double col[MMAX];
    int singular = 0;
    for (int i = 0; i < m; ++i) piv[i] = i;
    for (int j = 0; j < m; ++j) {
        for (int i = 0; i < m; ++i) col[i] = LU[i * m + j];
        for (int i = 0; i < m; ++i) {
            int kmax = i < j ? i : j;
            double s = 0.0;
            for (int k = 0; k < kmax; ++k) s += LU[i * m + k] * col[k];
            col[i] -= s;
            LU[i * m + j] = col[i];
        }
        int p = j;
        for (int i = j + 1; i < m; ++i)
            if (fabs(col[i]) > fabs(col[p])) p = i;
        if (p != j) {
            for (int k = 0; k < m; ++k) { double t = LU[p * m + k]; LU[p * m + k] = LU[j * m + k]; LU[j * m + k] = t; }
            int t = piv[p]; piv[p] = piv[j]; piv[j] = t;
        }
        if (LU[j * m + j] != 0.0) {
            for (int i = j + 1; i < m; ++i) LU[i * m + j] /= LU[j * m + j];
        } else {
            singular = 1;
        }
    }
    return singular;
}
static void lu_solve(int m, const double* LU, const int* piv, int nrhs, const double* Bm, double* X)
{
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < nrhs; ++j) X[i * nrhs + j] = Bm[piv[i] * nrhs + j];
    for (int k = 0; k < m; ++k)
        for (int i = k + 1; i < m; ++i)
            for (int j = 0; j < nrhs; ++j) X[i * nrhs + j] -= X[k * nrhs + j] * LU[i * m + k];
    for (int k = m - 1; k >= 0; --k) {
        for (int j = 0; j < nrhs; ++j) X[k * nrhs + j] /= LU[k * m + k];
        for (int i = 0; i < k; ++i)
            for (int j = 0; j < nrhs; ++j) X[i * nrhs + j] -= X[k * nrhs + j] * LU[i * m + k];
    }
}

/*
 * KartLQR.cs:17-128 for one problem.  Layout of the inputs = include/hk_abi.h (hk_lqng_solve_batch, one record).
 * Time-varying extension: Z,eta start from Q[T-1],q[T-1]; backward iteration t uses A[t],B[t],R[t]
 * in the coupled system and Q[t],q[t] in the update.  With time_varying == 0 this is the reference.
 */
int hk_oracle_lqng_solve(int N, int horizon, int time_varying,
                         const double* A, const double* B, const double* Q, const double* q, const double* R,
                         const double* x0,
                         double* u0, double* Pout, double* alphaout, double* traj)
{
    if (N < 1 || N > 4 || horizon < 0) return -1;
    const int n = XD * N, m = UD * N, T = horizon + 1;
    int status = 0;
    double Aj[NMAX * NMAX], Bj[4][NMAX * UD];
    double Z[4][NMAX * NMAX], eta[4][NMAX];
    double LHS[MMAX * MMAX], RHSMat[MMAX * NMAX], RHSVec[MMAX], P[MMAX * NMAX], alpha[MMAX];
    double F[NMAX * NMAX], beta[NMAX];
    double tmp1[NMAX * NMAX], tmp2[NMAX * NMAX], tmp3[NMAX * NMAX];
    double* Pall = (double*)malloc(sizeof(double) * T * m * n);
    double* aall = (double*)malloc(sizeof(double) * T * m);

    const size_t sA = (size_t)N * XD * XD, sB = (size_t)N * XD * UD, sQ = (size_t)N * n * n, sq = (size_t)N * n, sR = (size_t)N * UD * UD;
    const int tl = time_varying ? T - 1 : 0;
    for (int i = 0; i < N; ++i) {                                   /* Zs = Q, etas = q :62-63 */
        memcpy(Z[i], Q + tl * sQ + (size_t)i * n * n, sizeof(double) * n * n);
        memcpy(eta[i], q + tl * sq + (size_t)i * n, sizeof(double) * n);
    }
    for (int t = horizon; t >= 0; --t) {                            /* :64 */
        const int tt = time_varying ? t : 0;
        const double* At = A + tt * sA; const double* Bt = B + tt * sB;
        const double* Qt = Q + tt * sQ; const double* qt = q + tt * sq; const double* Rt = R + tt * sR;
        /* joint A = DiagonalStack :33-37 ; joint B_i = Stack(0.., B_i, ..0) :41-52 */
        memset(Aj, 0, sizeof(double) * n * n);
        for (int i = 0; i < N; ++i) {
            for (int r = 0; r < XD; ++r)
                for (int c = 0; c < XD; ++c) Aj[(XD * i + r) * n + XD * i + c] = At[i * XD * XD + r * XD + c];
            memset(Bj[i], 0, sizeof(double) * n * UD);
            for (int r = 0; r < XD; ++r)
                for (int c = 0; c < UD; ++c) Bj[i][(XD * i + r) * UD + c] = Bt[i * XD * UD + r * UD + c];
        }
        /* LHS :67-87.  col_i = vstack_j ( B_i^T (Z_i B_j) [+R_i if i==j] ); LHS = hstack_i col_i.
         * => block computed from (i,j) lands at ROW-block j, COLUMN-block i (quirk Q1). */
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) {
                matmul(n, n, UD, Z[i], Bj[j], tmp1);                /* Zs[i].Multiply(Bs[j]) */
                matmul_tn(n, UD, UD, Bj[i], tmp1, tmp2);            /* Bs[i].TransposeThisAndMultiply(..) */
                for (int a = 0; a < UD; ++a)
                    for (int b = 0; b < UD; ++b) {
                        double v = tmp2[a * UD + b];
                        if (i == j) v = Rt[i * UD * UD + a * UD + b] + v;   /* getRMatrix() + ... :78 */
                        LHS[(UD * j + a) * m + (UD * i + b)] = v;
                    }
            }
        /* RHSMat = vstack_i B_i^T (Z_i A), RHSVec = concat_i B_i^T eta_i :89-98 */
        for (int i = 0; i < N; ++i) {
            matmul(n, n, n, Z[i], Aj, tmp1);
            matmul_tn(n, UD, n, Bj[i], tmp1, tmp2);
            memcpy(RHSMat + (size_t)UD * i * n, tmp2, sizeof(double) * UD * n);
            matmul_tn(n, UD, 1, Bj[i], eta[i], tmp2);
            RHSVec[UD * i + 0] = tmp2[0]; RHSVec[UD * i + 1] = tmp2[1];
        }
        /* P = LHS.Solve(RHSMat); alpha = LHS.Solve(RHSVec) :104-105 */
        {
            double LU[MMAX * MMAX]; int piv[MMAX];
            memcpy(LU, LHS, sizeof(double) * m * m);
            status |= lu_factor(m, LU, piv);
            lu_solve(m, LU, piv, n, RHSMat, P);
            lu_solve(m, LU, piv, 1, RHSVec, alpha);
        }
        memcpy(Pall + (size_t)t * m * n, P, sizeof(double) * m * n);
        memcpy(aall + (size_t)t * m, alpha, sizeof(double) * m);
        /* F = A - sum_k B_k P_k ; beta = - sum_k B_k alpha_k :110-111 (Aggregate starts from a zero accumulator) */
        memset(tmp3, 0, sizeof(double) * n * n);
        memset(beta, 0, sizeof(double) * n);
        for (int k = 0; k < N; ++k) {
            matmul(n, UD, n, Bj[k], P + (size_t)UD * k * n, tmp1);
            for (int e = 0; e < n * n; ++e) tmp3[e] = tmp3[e] + tmp1[e];
            matmul(n, UD, 1, Bj[k], alpha + UD * k, tmp2);
            for (int e = 0; e < n; ++e) beta[e] = beta[e] - tmp2[e];
        }
        for (int e = 0; e < n * n; ++e) F[e] = Aj[e] - tmp3[e];
        /* updates :113-119, player order, eta uses the NEW Z_i (quirk Q2) */
        for (int i = 0; i < N; ++i) {
            const double* Pi = P + (size_t)UD * i * n;              /* SubMatrix(2i, 2, 0, n) */
            const double* Ri = Rt + i * UD * UD;
            double RP[UD * NMAX], PRP[NMAX * NMAX], ZF[NMAX * NMAX], FZF[NMAX * NMAX];
            matmul(UD, UD, n, Ri, Pi, RP);                          /* R_i P_i */
            matmul_tn(UD, n, n, Pi, RP, PRP);                       /* P_i^T (R_i P_i) */
            matmul(n, n, n, Z[i], F, ZF);                           /* Z_i F */
            matmul_tn(n, n, n, F, ZF, FZF);                         /* F^T (Z_i F) */
            for (int e = 0; e < n * n; ++e) Z[i][e] = (Qt[(size_t)i * n * n + e] + PRP[e]) + FZF[e];   /* :116 */
            double Ra[UD], PRa[NMAX], Zb[NMAX], FE[NMAX];
            matmul(UD, UD, 1, Ri, alpha + UD * i, Ra);              /* R_i alpha_i */
            matmul_tn(UD, n, 1, Pi, Ra, PRa);                       /* P_i^T (R_i alpha_i) */
            matmul(n, n, 1, Z[i], beta, Zb);                        /* Zs[i] (already updated) * beta */
            for (int e = 0; e < n; ++e) Zb[e] = eta[i][e] + Zb[e];
            matmul_tn(n, n, 1, F, Zb, FE);
            for (int e = 0; e < n; ++e) eta[i][e] = (qt[(size_t)i * n + e] + PRa[e]) + FE[e];          /* :117 */
        }
    }
    /* optimal_control = -P * initial - alpha with the t = 0 pair :121-126 (all players; reference keeps rows 0..1) */
    for (int r = 0; r < m; ++r) {
        double s = 0.0;
        for (int c = 0; c < n; ++c) s += -P[r * n + c] * x0[c];
        u0[r] = s - alpha[r];
    }
    if (Pout) memcpy(Pout, Pall, sizeof(double) * T * m * n);
    if (alphaout) memcpy(alphaout, aall, sizeof(double) * T * m);
    if (traj) {                                                     /* A.5 rollout */
        double x[NMAX], xn[NMAX], u[MMAX];
        memcpy(x, x0, sizeof(double) * n);
        memcpy(traj, x, sizeof(double) * n);
        for (int t = 0; t <= horizon; ++t) {
            const int tt = time_varying ? t : 0;
            const double* At = A + tt * sA; const double* Bt = B + tt * sB;
            const double* Pt = Pall + (size_t)t * m * n; const double* at = aall + (size_t)t * m;
            for (int r = 0; r < m; ++r) {
                double s = 0.0;
                for (int c = 0; c < n; ++c) s += -Pt[r * n + c] * x[c];
                u[r] = s - at[r];
            }
            for (int i = 0; i < N; ++i)
                for (int r = 0; r < XD; ++r) {
                    double s = 0.0;
                    for (int c = 0; c < XD; ++c) s += At[i * XD * XD + r * XD + c] * x[XD * i + c];
                    for (int c = 0; c < UD; ++c) s += Bt[i * XD * UD + r * UD + c] * u[UD * i + c];
                    xn[XD * i + r] = s;
                }
            memcpy(x, xn, sizeof(double) * n);
            memcpy(traj + (size_t)(t + 1) * n, x, sizeof(double) * n);
        }
    }
    free(Pall); free(aall);
    return status;
}

/* batch driver used as the CPU baseline: `threads` <= 1 runs serially, otherwise OpenMP static split */
int hk_oracle_lqng_solve_batch(int batch, int N, int horizon, int time_varying,
                               const double* A, const double* B, const double* Q, const double* q, const double* R,
                               const double* x0, double* u0, double* P, double* alpha, double* traj, int* status,
                               int threads)
{
    const int n = XD * N, m = UD * N, T = horizon + 1, Tm = time_varying ? T : 1;
    const size_t sA = (size_t)Tm * N * XD * XD, sB = (size_t)Tm * N * XD * UD, sQ = (size_t)Tm * N * n * n,
                 sq = (size_t)Tm * N * n, sR = (size_t)Tm * N * UD * UD;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int b = 0; b < batch; ++b) {
        int s = hk_oracle_lqng_solve(N, horizon, time_varying, A + b * sA, B + b * sB, Q + b * sQ, q + b * sq, R + b * sR,
                                     x0 + (size_t)b * n, u0 + (size_t)b * m,
                                     P ? P + (size_t)b * T * m * n : 0, alpha ? alpha + (size_t)b * T * m : 0,
                                     traj ? traj + (size_t)b * (T + 1) * n : 0);
        if (status) status[b] = s;
    }
    return 0;
}
