"""GPU parity of the LQNG kernels against the CPU oracle, through the C-ABI (include/hk_abi.h).
Tolerance: 1e-9 relative (BASELINE.json north_star), metric of SURVEY.md A.7 — see conftest.rel_err."""
import numpy as np
import pytest

from conftest import rel_err
from hierarchicalkarting_b200 import lqr, scenarios as S

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _check(got, ref, full=True):
    for b in range(ref["u0"].shape[0]):
        assert rel_err(got["u0"][b], ref["u0"][b]) <= TOL
        if full:
            for k in ("P", "alpha", "traj"):
                assert rel_err(got[k][b], ref[k][b]) <= TOL, (k, b)
    assert np.array_equal(got["status"], ref["status"])


@pytest.mark.parametrize("N,track", [(1, S.OVAL), (2, S.OVAL), (3, S.COMPLEX), (4, S.COMPLEX)])
@pytest.mark.parametrize("horizon", [3, 0, 7])
def test_parity_time_invariant(hk, oracle, N, track, horizon):
    p = S.make_problems(track, 257, N, seed=100 + N)          # ragged: not a multiple of any group/warp size
    A, B, Q, q, R, x0 = S.assemble_dense(p)
    ref = oracle.lqng_solve_batch(A, B, Q, q, R, x0, horizon)
    got = lqr.solve_batch(A, B, Q, q, R, x0, horizon)
    _check(got, ref)
    got_u = lqr.solve_batch(A, B, Q, q, R, x0, horizon, full=False)   # reference output only (u0)
    _check(got_u, ref, full=False)


@pytest.mark.parametrize("N", [1, 2, 3, 4])
def test_parity_time_varying(hk, oracle, N):
    rng = np.random.default_rng(5)
    horizon, batch = 3, 64
    T, n = horizon + 1, 4 * N
    p = S.make_problems(S.OVAL if N < 3 else S.COMPLEX, batch, N, seed=11)
    A, B, Q, q, R, x0 = S.assemble_dense(p)
    tv = lambda a, s: np.ascontiguousarray(np.repeat(a[:, None], T, axis=1) * (1.0 + s * rng.standard_normal((batch, T) + a.shape[1:])))
    At, Bt, Rt, qt = tv(A, 0.01), tv(B, 0.05), tv(R, 0.05), tv(q, 0.05)
    Qt = tv(Q, 0.05)
    Qt = 0.5 * (Qt + np.swapaxes(Qt, -1, -2))
    ref = oracle.lqng_solve_batch(At, Bt, Qt, qt, Rt, x0, horizon, time_varying=True)
    got = lqr.solve_batch(At, Bt, Qt, qt, Rt, x0, horizon, time_varying=True)
    _check(got, ref)
    got_u = lqr.solve_batch(At, Bt, Qt, qt, Rt, x0, horizon, time_varying=True, full=False)     # u0 only
    _check(got_u, ref, full=False)
    # stages that break the fast path of the 2-kart DMMA kernel (whole horizon staged by TMA): a non-symmetric Q_t at an inner
    # stage (shared-memory fallback inside the kernel), an R_t that forces row exchanges at one stage (pivot pass)
    for b in range(0, batch, 5):
        Qt[b, 1, 0] += 0.05 * rng.standard_normal((n, n))
    for b in range(1, batch, 5):
        c = 0.2 + 0.1 * rng.random()
        Rt[b, 2, :] = np.array([[0.05 * c, c], [c, 0.03 * c]])
    ref = oracle.lqng_solve_batch(At, Bt, Qt, qt, Rt, x0, horizon, time_varying=True)
    got = lqr.solve_batch(At, Bt, Qt, qt, Rt, x0, horizon, time_varying=True)
    _check(got, ref)
    got_u = lqr.solve_batch(At, Bt, Qt, qt, Rt, x0, horizon, time_varying=True, full=False)
    _check(got_u, ref, full=False)


def test_general_dense_blocks_and_nonsymmetric_q(hk, oracle):
    """The ABI takes any per-player A_i (4x4), B_i (4x2), R_i (2x2) and Q_i — not only the bicycle structure."""
    rng = np.random.default_rng(9)
    for N in (2, 3, 4):
        batch, n = 33, 4 * N
        A = np.eye(4) + 0.05 * rng.standard_normal((batch, N, 4, 4))
        B = 0.05 * rng.standard_normal((batch, N, 4, 2))
        Q = 0.3 * rng.standard_normal((batch, N, n, n))                 # deliberately NOT symmetric
        q = rng.standard_normal((batch, N, n))
        R = np.eye(2) * 0.2 + 0.02 * rng.standard_normal((batch, N, 2, 2))
        x0 = rng.standard_normal((batch, n))
        ref = oracle.lqng_solve_batch(A, B, Q, q, R, x0, 3)
        got = lqr.solve_batch(A, B, Q, q, R, x0, 3)
        _check(got, ref)


@pytest.mark.parametrize("N", [2, 3, 4])
def test_pivoting_and_singular_status(hk, oracle, N):
    """LHS whose natural pivot order is wrong (tiny R, large coupling) and an exactly singular LHS (R = 0, B = 0).
    N = 3 runs in the 4-kart kernel's frame with a decoupled dummy player: its identity rows must never be taken as pivots."""
    rng = np.random.default_rng(3)
    n, batch = 4 * N, 16
    A = np.tile(np.eye(4), (batch, N, 1, 1))
    B = rng.standard_normal((batch, N, 4, 2))
    Q = rng.standard_normal((batch, N, n, n)); Q = Q + np.swapaxes(Q, -1, -2)
    q = rng.standard_normal((batch, N, n))
    R = np.zeros((batch, N, 2, 2)); R[..., 0, 1] = 1e-3; R[..., 1, 0] = 1e-3      # zero diagonal forces row exchanges
    x0 = rng.standard_normal((batch, n))
    ref = oracle.lqng_solve_batch(A, B, Q, q, R, x0, 2)
    got = lqr.solve_batch(A, B, Q, q, R, x0, 2)
    assert np.array_equal(got["status"], ref["status"])
    for b in range(batch):
        for k in ("u0", "P", "alpha"):
            assert rel_err(got[k][b], ref[k][b]) <= (1e-7 if N == 2 else 1e-5), (k, b)   # conditioning of these systems is ~1e3-1e6
    Bz = np.zeros_like(B); Rz = np.zeros_like(R)
    ref = oracle.lqng_solve_batch(A, Bz, Q, q, Rz, x0, 1)
    got = lqr.solve_batch(A, Bz, Q, q, Rz, x0, 1)
    assert np.all(ref["status"] == 1) and np.array_equal(got["status"], ref["status"])


def test_pivot_pass_of_the_throughput_kernel(hk, oracle):
    """u0-only 2-kart batches take the DMMA kernel; problems whose coupled system needs row exchanges are re-solved by the
    same warp with partial pivoting (hk_lqng_mma2p.cuh), zero pivots go to the shared-memory algorithm.  Mixed batch."""
    rng = np.random.default_rng(21)
    batch = 301
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(S.OVAL, batch, 2, seed=5))
    kind = np.arange(batch) % 7
    for b in np.nonzero(kind == 3)[0]:                   # R_i with a dominant off-diagonal: every step exchanges rows
        c = 0.2 + 0.1 * rng.random()
        R[b] = np.array([[0.05 * c, c], [c, 0.03 * c]])
    for b in np.nonzero(kind == 5)[0]:                   # coupling larger than R: exchanges across the two players' rows
        B[b] = 3.0 * rng.standard_normal((2, 4, 2))
        Qb = rng.standard_normal((2, 8, 8))
        Q[b] = Qb + np.swapaxes(Qb, -1, -2)
        R[b] = np.eye(2) * 1e-2
    sing = np.nonzero(kind == 6)[0][:5]                  # exactly singular: status 1 like the oracle
    B[sing] = 0.0
    R[sing] = 0.0
    ref = oracle.lqng_solve_batch(A, B, Q, q, R, x0, 3, full=False)
    got = lqr.solve_batch(A, B, Q, q, R, x0, 3, full=False)
    assert np.array_equal(got["status"], ref["status"]) and ref["status"].sum() == 5
    for b in range(batch):
        if ref["status"][b]:
            continue
        tol = 1e-7 if kind[b] == 5 else TOL               # kind 5: condition numbers up to ~1e5
        assert rel_err(got["u0"][b], ref["u0"][b]) <= tol, (b, kind[b])


@pytest.mark.gpu
def test_full_outputs_of_the_throughput_kernel(hk, oracle):
    """Gains, offsets and the closed-loop rollout of a time-invariant 2-kart batch come from the same DMMA kernel (FULL mode of
    hk_lqng_mma2p.cuh): mixed batch with row-exchange problems (pivot pass), non-symmetric Q (shared-memory fallback inside the
    kernel), singular systems; with every subset of the optional outputs requested."""
    rng = np.random.default_rng(22)
    batch = 257
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(S.OVAL, batch, 2, seed=6))
    kind = np.arange(batch) % 7
    for b in np.nonzero(kind == 3)[0]:
        c = 0.2 + 0.1 * rng.random()
        R[b] = np.array([[0.05 * c, c], [c, 0.03 * c]])
    for b in np.nonzero(kind == 4)[0]:
        Q[b] += 0.05 * rng.standard_normal((2, 8, 8))      # non-symmetric Q_i
    sing = np.nonzero(kind == 6)[0][:4]
    B[sing] = 0.0
    R[sing] = 0.0
    ref = oracle.lqng_solve_batch(A, B, Q, q, R, x0, 3)
    got = lqr.solve_batch(A, B, Q, q, R, x0, 3)
    assert np.array_equal(got["status"], ref["status"]) and ref["status"].sum() == 4
    ok = ref["status"] == 0
    for k in ("u0", "P", "alpha", "traj"):
        for b in np.nonzero(ok)[0]:
            assert rel_err(got[k][b], ref[k][b]) <= TOL, (k, b, kind[b])
    # more than 8 steps: the rollout re-reads its gains from the output buffers instead of the warp's shared-memory copy
    ref9 = oracle.lqng_solve_batch(A[:40], B[:40], Q[:40], q[:40], R[:40], x0[:40], 9)
    got9 = lqr.solve_batch(A[:40], B[:40], Q[:40], q[:40], R[:40], x0[:40], 9)
    assert np.array_equal(got9["status"], ref9["status"])
    for k in ("u0", "P", "alpha", "traj"):
        for b in np.nonzero(ref9["status"] == 0)[0]:
            assert rel_err(got9[k][b], ref9[k][b]) <= TOL, (k, b)
    # subsets of the optional outputs through the device-pointer entry (the launcher lends scratch for what the rollout re-reads)
    import torch
    from hierarchicalkarting_b200 import abi
    lib = hk
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (A, B, Q, q, R, x0)]
    for want in (("P",), ("alpha",), ("traj",), ("P", "traj")):
        u0 = torch.zeros((batch, 4), dtype=torch.float64, device=dev)
        st = torch.zeros(batch, dtype=torch.int32, device=dev)
        out = {"P": torch.zeros((batch, 4, 4, 8), dtype=torch.float64, device=dev), "alpha": torch.zeros((batch, 4, 4), dtype=torch.float64, device=dev),
               "traj": torch.zeros((batch, 5, 8), dtype=torch.float64, device=dev)}
        ptr = {k: (out[k].data_ptr() if k in want else None) for k in out}
        abi.check(lib.hk_lqng_solve_batch_device(batch, 2, 3, 0, *[t.data_ptr() for t in d], u0.data_ptr(), ptr["P"], ptr["alpha"], ptr["traj"],
                                                 st.data_ptr(), None))
        torch.cuda.synchronize()
        assert np.array_equal(st.cpu().numpy(), ref["status"])
        for k in ("u0",) + want:
            g = (u0 if k == "u0" else out[k]).cpu().numpy()
            for b in np.nonzero(ok)[0]:
                assert rel_err(g[b], ref[k][b]) <= TOL, (want, k, b)


def test_golden_fixture(hk):
    g = np.load("tests/golden/lqng_golden.npz")
    for N in (2, 4):
        got = lqr.solve_batch(*(g[f"N{N}_{k}"] for k in ("A", "B", "Q", "q", "R", "x0")), 3)
        for k in ("u0", "P", "alpha", "traj"):
            for b in range(got["u0"].shape[0]):
                assert rel_err(got[k][b], g[f"N{N}_{k}_out"][b]) <= TOL


def test_reference_api_single_solve(hk, oracle):
    """KartLQR.solveFeedbackLQR drop-in on BASELINE config 1 (one 2-kart Oval problem) and the SURVEY Appendix D smoke value."""
    p = S.config1()
    A, B, Q, q, R, x0 = S.assemble_dense(p)
    dyn = [lqr.LinearizedBicycle(p["dt"], p["x0"][0, i]) for i in range(2)]

    class Cost(lqr.KartLQRCosts):
        def __init__(self, i): self.i = i
        def getQMatrix(self): return Q[0, self.i]
        def getQVec(self): return q[0, self.i]
        def getRMatrix(self): return R[0, self.i]
    u = lqr.KartLQR.solveFeedbackLQR(dyn, [Cost(0), Cost(1)], [p["x0"][0, 0], p["x0"][0, 1]], 3)
    ref = oracle.lqng_solve_batch(A, B, Q, q, R, x0, 3)
    assert u.shape == (2,) and rel_err(u, ref["u0"][0, :2]) <= TOL


def test_edge_cases(hk):
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(S.OVAL, 1, 2, seed=1))
    out = lqr.solve_batch(A[:0], B[:0], Q[:0], q[:0], R[:0], x0[:0], 3)     # empty batch
    assert out["u0"].shape == (0, 4)
    with pytest.raises(ValueError):
        lqr.solve_batch(A, B, Q[:, :, :7], q, R, x0, 3)                      # dimension mismatch -> ArgumentException analogue
    from hierarchicalkarting_b200 import abi
    lib = abi.load_library()
    u0 = np.zeros(4)
    rc = lib.hk_lqng_solve_batch(1, 5, 3, 0, abi.dptr(A), abi.dptr(B), abi.dptr(Q), abi.dptr(q), abi.dptr(R), abi.dptr(x0), abi.dptr(u0),
                                 None, None, None, None)
    assert rc == abi.HK_ERR_INVALID_ARGUMENT and b"n_players" in lib.hk_last_error()


def test_full_size_properties(hk, oracle):
    """BASELINE config 2 size (65,536 problems): determinism, batch-permutation equivariance, duplicate problems give
    identical answers, and EVERY problem checked against the oracle (all host threads: ~50 ms)."""
    p = S.config2(65536)
    A, B, Q, q, R, x0 = S.assemble_dense(p)
    a = lqr.solve_batch(A, B, Q, q, R, x0, 3, full=False)
    b = lqr.solve_batch(A, B, Q, q, R, x0, 3, full=False)
    assert np.array_equal(a["u0"], b["u0"])
    perm = np.random.default_rng(0).permutation(65536)
    c = lqr.solve_batch(A[perm], B[perm], Q[perm], q[perm], R[perm], x0[perm], 3, full=False)
    assert np.array_equal(c["u0"], a["u0"][perm])
    import os
    ref = oracle.lqng_solve_batch(A, B, Q, q, R, x0, 3, full=False, threads=os.cpu_count() or 1)
    scale = np.maximum(np.abs(ref["u0"]), np.max(np.abs(ref["u0"]), axis=1, keepdims=True))
    err = np.abs(a["u0"] - ref["u0"]) / np.where(scale == 0, 1.0, scale)                 # SURVEY.md A.7 metric, per problem
    assert float(err.max()) <= TOL, (float(err.max()), int(np.argmax(err.max(axis=1))))
    assert np.all(a["status"] == 0) and np.all(ref["status"] == 0)


def test_full_size_properties_4kart(hk, oracle):
    """BASELINE config 3 shape (4-kart 2v2 Complex), 131,072 problems through the DMMA kernel: determinism, batch-permutation
    equivariance, and a strided sample against the oracle; all outputs on a smaller slice."""
    n = 131072
    p = S.make_problems(S.COMPLEX, n, 4, seed=20260002)
    A, B, Q, q, R, x0 = S.assemble_dense(p)
    a = lqr.solve_batch(A, B, Q, q, R, x0, 3, full=False)
    b = lqr.solve_batch(A, B, Q, q, R, x0, 3, full=False)
    assert np.array_equal(a["u0"], b["u0"]) and np.all(a["status"] == 0)
    perm = np.random.default_rng(1).permutation(n)[:32768]
    c = lqr.solve_batch(A[perm], B[perm], Q[perm], q[perm], R[perm], x0[perm], 3, full=False)
    assert np.array_equal(c["u0"], a["u0"][perm])
    import os
    every = oracle.lqng_solve_batch(A, B, Q, q, R, x0, 3, full=False, threads=os.cpu_count() or 1)     # all 131,072 problems
    scale = np.maximum(np.abs(every["u0"]), np.max(np.abs(every["u0"]), axis=1, keepdims=True))
    err = np.abs(a["u0"] - every["u0"]) / np.where(scale == 0, 1.0, scale)
    assert float(err.max()) <= TOL, (float(err.max()), int(np.argmax(err.max(axis=1))))
    idx = np.arange(0, n, 509)
    ref = oracle.lqng_solve_batch(A[idx], B[idx], Q[idx], q[idx], R[idx], x0[idx], 3)
    full = lqr.solve_batch(A[idx], B[idx], Q[idx], q[idx], R[idx], x0[idx], 3)
    _check(full, ref)


def test_assemble_on_device(hk, oracle):
    for N, track in ((2, S.OVAL), (4, S.COMPLEX), (1, S.OVAL), (3, S.COMPLEX)):
        p = S.make_problems(track, 300, N, seed=77)
        ref = oracle.lqng_solve_batch(*S.assemble_dense(p), 3, full=False)
        got = lqr.assemble_solve_batch(p, 3)
        for b in range(300):
            assert rel_err(got["u0"][b], ref["u0"][b]) <= TOL


def test_packed_records_equal_the_seven_array_call(hk, oracle):
    """hk_lqng_assemble_solve_packed (one interleaved record per problem: one H2D copy per chunk, one TMA bulk copy per problem in the
    2-kart kernel) returns bit for bit what hk_lqng_assemble_solve_batch returns, for every player count, ragged sizes and the multi-chunk
    pipeline (131,075 two-kart problems: 9 chunks through the ring of four), and agrees with the oracle."""
    for N, track, batch in ((2, S.OVAL, 301), (4, S.COMPLEX, 300), (1, S.OVAL, 77), (3, S.COMPLEX, 130), (2, S.OVAL, 131075)):
        p = S.make_problems(track, batch, N, seed=78)
        a = lqr.assemble_solve_batch(p, 3)
        rec = lqr.pack_records(p)
        assert rec.shape == (batch, 13 * N + 9 * N * (N - 1))
        b = lqr.assemble_solve_packed(rec, N, 3, p["dt"])
        assert np.array_equal(a["u0"], b["u0"]) and np.array_equal(a["status"], b["status"]), (N, batch)
        idx = np.arange(0, batch, max(1, batch // 200))
        sub = {k: (v[idx] if isinstance(v, np.ndarray) and v.shape[:1] == (batch,) else v) for k, v in p.items()}
        ref = oracle.lqng_solve_batch(*S.assemble_dense(sub), 3, full=False)
        for k, i in enumerate(idx):
            assert rel_err(b["u0"][i], ref["u0"][k]) <= TOL


def test_assemble_pipeline_chunks_and_ring(hk, oracle):
    """hk_lqng_assemble_solve_batch at sizes that exercise its chunk pipeline: 2 chunks (20,001 problems, odd), 8 chunks over a ring of
    4 chunk buffers with reuse waits (70,001 problems), pageable and pinned caller buffers; every problem against the dense DMMA path
    (itself checked against the oracle), a sample against the oracle directly."""
    import torch
    for batch in (20001, 70001):
        p = S.make_problems(S.OVAL, batch, 2, seed=78)
        dense = lqr.solve_batch(*S.assemble_dense(p), 3, full=False)
        got = lqr.assemble_solve_batch(p, 3)                                     # pageable numpy buffers
        assert np.array_equal(got["status"], dense["status"])
        scale = np.maximum(np.abs(dense["u0"]), np.abs(dense["u0"]).max(axis=1, keepdims=True))
        assert np.max(np.abs(got["u0"] - dense["u0"]) / scale) <= TOL
        idx = np.linspace(0, batch - 1, 257).astype(int)
        ref = oracle.lqng_solve_batch(*[a[idx] for a in S.assemble_dense(p)], 3, full=False)
        for j, b in enumerate(idx):
            assert rel_err(got["u0"][b], ref["u0"][j]) <= TOL
        keys = ("x0", "target", "tw", "cw", "aw", "otgt", "otw")                 # pinned caller buffers (the batched-copy path)
        pin = [torch.from_numpy(np.ascontiguousarray(p[k], dtype=np.float64)).pin_memory().numpy() for k in keys]
        u0 = torch.empty((batch, 4), dtype=torch.float64).pin_memory().numpy()
        st = torch.empty(batch, dtype=torch.int32).pin_memory().numpy()
        from hierarchicalkarting_b200 import abi
        for rep in range(2):
            u0[:] = 0
            abi.check(hk.hk_lqng_assemble_solve_batch(batch, 2, 3, float(p["dt"]), *[abi.dptr(a) for a in pin], abi.dptr(u0), abi.iptr(st)))
            assert np.array_equal(u0, got["u0"]) and np.array_equal(st, got["status"])


def test_reentrant_from_several_host_threads(hk, oracle):
    """The reference calls the solver from the Unity main thread and the MCTS from one background thread per agent
    (HierarchicalKartAgent.cs:246-283): concurrent host threads get their own stream / scratch and must not disturb each other
    (large batches through the persistent kernel with its shared counter pool, small calls, rollouts)."""
    import threading
    from hierarchicalkarting_b200 import mcts as M, tracks
    probs = [S.assemble_dense(S.make_problems(S.OVAL, 20000 + 37 * k, 2, seed=500 + k)) for k in range(4)]
    want = [lqr.solve_batch(*p, 3, full=False)["u0"] for p in probs]
    G = M.Game(tracks.COMPLEX, 2, 2)
    leaf = tracks.root_state(tracks.COMPLEX, 5, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 40])
    want_r = G.rollouts(leaf, 50000, seed=9)
    errors = []

    def lq_worker(k):
        try:
            for rep in range(6):
                got = lqr.solve_batch(*probs[k], 3, full=False)["u0"]
                if not np.array_equal(got, want[k]):
                    errors.append(("lqng", k, rep))
                one = lqr.solve_batch(*(a[:1] for a in probs[k]), 3, full=False)["u0"]
                if not np.array_equal(one[0], want[k][0]):
                    errors.append(("lqng-one", k, rep))
        except Exception as e:                                   # noqa: BLE001
            errors.append(("exc", k, repr(e)))

    def mcts_worker():
        try:
            for rep in range(6):
                got = G.rollouts(leaf, 50000, seed=9)
                if not (np.array_equal(got["visit"], want_r["visit"]) and np.allclose(got["reward_sum"], want_r["reward_sum"], rtol=1e-12, atol=1e-9)):
                    errors.append(("mcts", rep))
        except Exception as e:                                   # noqa: BLE001
            errors.append(("exc-mcts", repr(e)))

    threads = [threading.Thread(target=lq_worker, args=(k,)) for k in range(4)] + [threading.Thread(target=mcts_worker)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
    ref = oracle.lqng_solve_batch(*(a[:64] for a in probs[0]), 3, full=False)["u0"]
    for b in range(64):
        assert rel_err(want[0][b], ref[b]) <= TOL
