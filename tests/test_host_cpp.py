"""The compiled (C++) host mirror of the reference API — host/KartLQR.hpp, host/KartMCTS.hpp — builds against
include/hk_abi.h (CPU test) and, on a GPU box, solves BASELINE config 1 through KartLQR::solveFeedbackLQR to 1e-9 of the
oracle and runs a short KartMCTS::constructSearchTree."""
import os
import subprocess

import numpy as np
import pytest

from hierarchicalkarting_b200 import abi, scenarios as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "test_host")


def _build():
    import __graft_entry__ as g
    g.build()
    libdir = os.path.dirname(abi.LIB_PATH)
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", EXE, os.path.join(ROOT, "host", "test_host.cpp"),
                           "-L" + libdir, "-lhk_b200", "-Wl,-rpath," + libdir])


def test_host_mirror_compiles_against_abi():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_host_mirror_runs(hk, oracle, tmp_path):
    _build()
    p = S.config1()
    ref = oracle.lqng_solve_batch(*S.assemble_dense(p), 3, full=False)
    rows = [repr(float(p["dt"]))]
    for i in range(2):
        for k in ("x0", "target", "tw"):
            rows.append(" ".join(repr(float(v)) for v in p[k][0, i]))
        rows.append(" ".join(repr(float(v)) for v in (p["cw"][0, i], p["aw"][0, i, 0, 0], p["aw"][0, i, 0, 1])))
        rows.append(" ".join(repr(float(v)) for v in p["otgt"][0, i, 0]))
        rows.append(" ".join(repr(float(v)) for v in p["otw"][0, i, 0]))
    rows.append(" ".join(repr(float(v)) for v in ref["u0"][0, :2]))
    f = tmp_path / "problem.txt"
    f.write_text("\n".join(rows) + "\n")
    out = subprocess.run([EXE, str(f)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "HOST_OK" in out.stdout, out.stdout + out.stderr
