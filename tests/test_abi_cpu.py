"""CPU tests (no GPU): the C-ABI shared library loads, exports every symbol include/hk_abi.h declares, validates its
arguments on the host, and refuses to compute without a CUDA device (there is no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from hierarchicalkarting_b200 import abi, scenarios as S

HEADER = open("include/hk_abi.h").read()


def _declared():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(hk_[a-z0-9_]+)\s*\(", body)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    lib = abi.load_library()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hk_abi.h but not exported"
        assert n in abi.PROTOTYPES, f"{n} has no ctypes prototype"
    syms = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True).stdout
    for n in names:
        assert re.search(rf"\bT {n}\b", syms)
    assert lib.hk_abi_version() == abi.HK_ABI_VERSION


def test_only_sm100a_code_in_library():
    out = subprocess.run(["cuobjdump", "-lelf", abi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts_match_header():
    assert C.sizeof(abi.hk_section) == 24 and C.sizeof(abi.hk_kart) == 28 and C.sizeof(abi.hk_game_params) == 32
    assert C.sizeof(abi.hk_kart_state) == 40 and C.sizeof(abi.hk_action) == 12 and C.sizeof(abi.hk_game_state) == 16 + 4 * 40
    assert C.sizeof(abi.hk_race_kart) == 64 and C.sizeof(abi.hk_race_plan) == 936 and C.sizeof(abi.hk_race_params) == 56
    assert C.sizeof(abi.hk_race_mcts_params) == 40 and abi.MCTS_NODE_DTYPE.itemsize == 32
    assert re.search(r"#define HK_MAX_SECTIONS 64\b", HEADER) and abi.HK_MAX_SECTIONS == 64
    for name, val in (("HK_MAX_PLAYERS", 4), ("HK_MAX_ACTIONS", 36), ("HK_MAX_PLIES", 64), ("HK_MAX_KARTS", 4), ("HK_MAX_HORIZON", 31)):
        assert re.search(rf"#define {name} {val}\b", HEADER) and getattr(abi, name) == val


def _parse_header_structs():
    """{struct name: [(field, C type, array length or None)]} of every `typedef struct hk_x { ... } hk_x;` in include/hk_abi.h."""
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    defines = {k: int(v) for k, v in re.findall(r"#define\s+(HK_[A-Z_]+)\s+(\d+)", body)}
    out = {}
    for name, fields in re.findall(r"typedef struct (hk_[a-z_]+)\s*\{(.*?)\}\s*\1\s*;", body, flags=re.S):
        rows = []
        for decl in fields.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"([a-z0-9_]+(?:\s+[a-z0-9_]+)*?)\s+(.+)$", decl)
            ctype, names = m.group(1), m.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                am = re.match(r"([A-Za-z0-9_]+)\[([A-Z0-9_]+)\]$", nm)
                if am:
                    ln = am.group(2)
                    rows.append((am.group(1), ctype, defines[ln] if ln in defines else int(ln)))
                else:
                    rows.append((nm, ctype, None))
        out[name] = rows
    return out


_CT = {"float": C.c_float, "double": C.c_double, "int32_t": C.c_int32, "int8_t": C.c_int8, "uint8_t": C.c_uint8, "uint64_t": C.c_uint64}


@pytest.mark.parametrize("which", ["product", "oracle"])
def test_struct_definitions_follow_the_header_field_by_field(which):
    """Both Python restatements of the ABI's plain-data structs — the product's (hierarchicalkarting_b200/abi.py) and the oracle's OWN
    (oracle/structs.py, written separately so that a layout bug is not common-mode) — are checked against the header text: field
    names, order, C types, array lengths, and therefore offsets and sizes."""
    if which == "product":
        mod = abi
    else:
        from oracle import structs as mod
    parsed = _parse_header_structs()
    assert len(parsed) >= 11
    for name, rows in parsed.items():
        cls = getattr(mod, name, None)
        assert cls is not None, f"{which}: no definition of {name}"
        got = []
        for fname, ftype in cls._fields_:
            if issubclass(ftype, C.Array):
                got.append((fname, ftype._type_, ftype._length_))
            else:
                got.append((fname, ftype, None))
        want = []
        for fname, ctype, ln in rows:
            t = _CT.get(ctype) or getattr(mod, ctype)
            want.append((fname, t, ln))
        assert [(a, c) for a, _, c in got] == [(a, c) for a, _, c in want], name
        for (fn, tg, _), (_, tw, _) in zip(got, want):
            assert tg is tw or (C.sizeof(tg) == C.sizeof(tw) and tg.__name__ == tw.__name__), (name, fn, tg, tw)


def test_host_side_argument_validation_needs_no_gpu():
    lib = abi.load_library()
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(S.OVAL, 2, 2, seed=1))
    u0 = np.zeros((2, 4))
    args = [abi.dptr(a) for a in (A, B, Q, q, R, x0)] + [abi.dptr(u0), None, None, None, None]
    assert lib.hk_lqng_solve_batch(2, 0, 3, 0, *args) == abi.HK_ERR_INVALID_ARGUMENT
    assert lib.hk_lqng_solve_batch(2, 2, 99, 0, *args) == abi.HK_ERR_INVALID_ARGUMENT
    assert lib.hk_lqng_solve_batch(-1, 2, 3, 0, *args) == abi.HK_ERR_INVALID_ARGUMENT
    assert b"batch" in lib.hk_last_error()
    assert lib.hk_lqng_solve_batch(0, 2, 3, 0, *args) == abi.HK_OK            # empty batch is a no-op, even without a device
    # closed-loop entries: NULL handles / bad sizes are rejected before any device work; empty batches are no-ops
    from hierarchicalkarting_b200 import race as R
    prm = R.race_params(S.OVAL)
    karts, plans = R.start_grid(S.OVAL, 2, seed=1)
    u = np.zeros((2, 2, 2))
    assert lib.hk_race_run(None, C.byref(prm), 2, 0, 10, abi.vptr(karts), abi.vptr(plans), abi.dptr(u), None) == abi.HK_ERR_INVALID_ARGUMENT
    assert b"hk_race_run" in lib.hk_last_error()
    assert lib.hk_race_step(None, C.byref(prm), 4, 0, abi.dptr(u), abi.vptr(karts), abi.vptr(plans)) == abi.HK_ERR_INVALID_ARGUMENT
    sec, trig, fwd, lane = R.geometry(S.OVAL)
    h = C.c_void_p()
    assert lib.hk_track_create(sec, abi.dptr(trig), abi.dptr(fwd), abi.dptr(lane), 65, C.byref(h)) == abi.HK_ERR_INVALID_ARGUMENT
    assert lib.hk_track_create(sec, None, abi.dptr(fwd), abi.dptr(lane), 24, C.byref(h)) == abi.HK_ERR_INVALID_ARGUMENT
    cdf = np.zeros(4, dtype=np.uint32)
    assert lib.hk_policy_cdf(0, cdf.ctypes.data_as(C.POINTER(C.c_uint32))) == abi.HK_ERR_INVALID_ARGUMENT
    assert lib.hk_policy_cdf(4, cdf.ctypes.data_as(C.POINTER(C.c_uint32))) == abi.HK_OK and cdf[-1] == 0xFFFFFFFF


def test_no_cpu_fallback():
    """Without a CUDA device the compute entries fail loudly with HK_ERR_NO_DEVICE."""
    lib = abi.load_library()
    if lib.hk_device_count() > 0:
        pytest.skip("a CUDA device is present")
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(S.OVAL, 2, 2, seed=1))
    u0 = np.zeros((2, 4))
    rc = lib.hk_lqng_solve_batch(2, 2, 3, 0, *[abi.dptr(a) for a in (A, B, Q, q, R, x0)], abi.dptr(u0), None, None, None, None)
    assert rc == abi.HK_ERR_NO_DEVICE and b"no CPU fallback" in lib.hk_last_error()
    from hierarchicalkarting_b200 import mcts, tracks
    with pytest.raises(abi.HKError) as e:
        mcts.Game(tracks.OVAL, 2, 2)
    assert e.value.status == abi.HK_ERR_NO_DEVICE
    from hierarchicalkarting_b200 import race
    with pytest.raises(abi.HKError) as e:
        race.Races(tracks.OVAL)                                              # the headless race loop has no CPU path either
    assert e.value.status == abi.HK_ERR_NO_DEVICE


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package, include/ or host/ may reference it."""
    for root in ("hierarchicalkarting_b200", "include", "host", "csharp"):
        for dp, _, fs in os.walk(root):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cs", ".cpp")):
                    txt = open(os.path.join(dp, f)).read()
                    assert "hk_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, os.path.join(dp, f)
