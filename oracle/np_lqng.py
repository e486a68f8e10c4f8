"""Independent numpy restatement of KartLQR.solveFeedbackLQR (Assets/Karting/Scripts/AI/LQR/KartLQR.cs:17-128) —
CPU ORACLE cross-check (test infrastructure, NOT the product).  Written against the C# text line by line with numpy
matrices (np.vstack/np.hstack for Stack/Append, scipy-free np.linalg.solve = LAPACK LU with partial pivoting) so that
the C restatement in hk_oracle_lqng.c is pinned by a second, independently written implementation."""
import numpy as np


def solve_feedback_lqr(As, Bs_local, Qs, qs, Rs, initials, horizon):
    players = len(As)
    xdims = [a.shape[0] for a in As]
    udims = [b.shape[1] for b in Bs_local]
    totalX, totalU = sum(xdims), sum(udims)
    uidx, cur = {}, 0
    for i in range(players):
        uidx[i] = (cur, udims[i]); cur += udims[i]
    A = np.zeros((0, 0))
    for i in range(players):                                   # DiagonalStack :33-37
        A = np.block([[A, np.zeros((A.shape[0], xdims[i]))], [np.zeros((xdims[i], A.shape[1])), As[i]]])
    Bs = []
    for i in range(players):                                   # :41-52
        B = np.zeros((0, udims[i]))
        for j in range(players):
            B = np.vstack([B, Bs_local[j] if i == j else np.zeros((xdims[j], udims[i]))])
        Bs.append(B)
    initial = np.concatenate(initials)
    Zs = [Q.copy() for Q in Qs]
    etas = [q.copy() for q in qs]
    Ps, alphas = [None] * (horizon + 1), [None] * (horizon + 1)
    for t in range(horizon, -1, -1):
        LHS = np.zeros((totalU, 0))
        for i in range(players):
            col = np.zeros((0, udims[i]))
            for j in range(players):
                blk = Bs[i].T @ (Zs[i] @ Bs[j])
                if i == j:
                    blk = Rs[i] + blk
                col = np.vstack([col, blk])                    # col.Stack :78/:82
            LHS = np.hstack([LHS, col])                        # LHS.Append(col) :85
        RHSMat = Bs[0].T @ (Zs[0] @ A)
        RHSVec = np.zeros(totalU)
        for i in range(players):
            if i > 0:
                RHSMat = np.vstack([RHSMat, Bs[i].T @ (Zs[i] @ A)])
            RHSVec[uidx[i][0]:uidx[i][0] + uidx[i][1]] = Bs[i].T @ etas[i]
        P = np.linalg.solve(LHS, RHSMat)
        alpha = np.linalg.solve(LHS, RHSVec)
        Ps[t], alphas[t] = P, alpha
        F = A - sum(Bs[k] @ P[uidx[k][0]:uidx[k][0] + uidx[k][1], :] for k in range(players))
        beta = -sum(Bs[k] @ alpha[uidx[k][0]:uidx[k][0] + uidx[k][1]] for k in range(players))
        for i in range(players):
            Pi = P[uidx[i][0]:uidx[i][0] + uidx[i][1], :]
            ai = alpha[uidx[i][0]:uidx[i][0] + uidx[i][1]]
            Zs[i] = Qs[i] + Pi.T @ (Rs[i] @ Pi) + F.T @ (Zs[i] @ F)                       # :116
            etas[i] = qs[i] + Pi.T @ (Rs[i] @ ai) + F.T @ (etas[i] + Zs[i] @ beta)       # :117 (new Z_i)
    P0 = P[uidx[0][0]:uidx[0][0] + uidx[0][1], :]
    a0 = alpha[uidx[0][0]:uidx[0][0] + uidx[0][1]]
    return -P0 @ initial - a0, np.stack(Ps), np.stack(alphas), (-P @ initial - alpha)
