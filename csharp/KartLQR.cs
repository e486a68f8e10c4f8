// KartLQR.cs — drop-in replacement of Assets/Karting/Scripts/AI/LQR/KartLQR.cs.
// Same namespace, class and signature (reference KartLQR.cs:17); the body only flattens the providers
// (KartLQRDynamics.getA/getB, KartLQRCosts.getQMatrix/getQVec/getRMatrix) into blittable arrays and calls the CUDA library.
// HierarchicalKartAgent.SolveLQR (HierarchicalKartAgent.cs:1201) compiles and runs unchanged.
using MathNet.Numerics.LinearAlgebra;
using System;
using System.Collections.Generic;
using KartGame.AI.Native;

namespace KartGame.AI.LQR
{
    public class KartLQR
    {
        public static Vector<double> solveFeedbackLQR(List<KartLQRDynamics> dynamics, List<KartLQRCosts> costs, List<Vector<double>> initials, int horizon)
        {
            int N = dynamics.Count, n = 4 * N;
            if (N < 1 || N > 4 || costs.Count != N || initials.Count != N) throw new ArgumentException("player count mismatch");
            var A = new double[N * 16]; var B = new double[N * 8]; var Q = new double[N * n * n]; var q = new double[N * n];
            var R = new double[N * 4]; var x0 = new double[n]; var u0 = new double[2 * N];
            for (int i = 0; i < N; i++)
            {
                if (dynamics[i].getXDim() != 4 || dynamics[i].getUDim() != 2) throw new ArgumentException("only 4-state / 2-control players");
                Copy(dynamics[i].getA(), 4, 4, A, i * 16);
                Copy(dynamics[i].getB(), 4, 2, B, i * 8);
                Copy(costs[i].getQMatrix(), n, n, Q, i * n * n);      // throws ArgumentException on a wrong size, like MathNet would later
                Copy(costs[i].getRMatrix(), 2, 2, R, i * 4);
                var qv = costs[i].getQVec();
                if (qv.Count != n || initials[i].Count != 4) throw new ArgumentException("dimension mismatch");
                for (int c = 0; c < n; c++) q[i * n + c] = qv[c];
                for (int c = 0; c < 4; c++) x0[4 * i + c] = initials[i][c];
            }
            HkNative.Check(HkNative.hk_lqng_solve_one(N, horizon, A, B, Q, q, R, x0, u0));
            return CreateVector.Dense(new[] { u0[0], u0[1] });         // player 0's first control (reference :121-127)
        }

        /// Batched form (new): `batch` independent games in the record layout of include/hk_abi.h; returns u0[batch][2N].
        public static double[] solveFeedbackLQRBatch(int batch, int players, int horizon, double[] A, double[] B, double[] Q, double[] q,
                                                     double[] R, double[] x0, int[] status = null)
        {
            var u0 = new double[batch * 2 * players];
            HkNative.Check(HkNative.hk_lqng_solve_batch(batch, players, horizon, 0, A, B, Q, q, R, x0, u0, null, null, null, status));
            return u0;
        }

        static void Copy(Matrix<double> m, int rows, int cols, double[] dst, int off)
        {
            if (m.RowCount != rows || m.ColumnCount != cols) throw new ArgumentException("dimension mismatch");
            for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) dst[off + r * cols + c] = m[r, c];
        }
    }
}
