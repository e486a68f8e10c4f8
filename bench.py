#!/usr/bin/env python
"""bench.py — BASELINE.json metric: LQNG solves/s (2-kart, reference horizon 3) on N B200s; MCTS rollouts/s beside it.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     the reference's CPU path (oracle port, all host threads)

A "step" is one pass of the hot path over one batch of synthetic problems: BASELINE config 2, 65,536 independent 2-kart
Oval problems per GPU (weak scaling: every rank solves its own seeded shard; no collective on the data path, one final
gather of per-rank summaries).  `value` times hk_lqng_solve_batch_device on HBM-resident inputs with CUDA events on the
launching stream; `e2e` times the host-pointer C-ABI call hk_lqng_solve_batch (pinned host buffers, H2D + D2H inside).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH = 65536
HORIZON = 3
FLOPS_PER_SOLVE = 41805            # dense count, SURVEY.md §8(d), 2-kart horizon 3
FLOPS_PER_SOLVE_4 = 556823         # same count for the 4-kart game
IN_BYTES_PER_SOLVE = 1664          # A,B per player + Q,q,R,x0 (per-player A/B form)
OUT_BYTES_PER_SOLVE = 32 + 4       # u0 for both players + status
N_INPUT_SETS = 4                   # rotated so that consecutive steps never hit the same 109 MB in the 126 MB L2
MCTS_ROLLOUTS = 1_000_000          # BASELINE config 4
RACES = 16384                      # BASELINE config 5
RACE_STEPS = 200


def _peaks():
    hbm, hbm_src = 6650.0, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm, hbm_src = float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        pass
    fp64, fp64_src = 37.2, "nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz"
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_peaks.json")) as f:
            j = json.load(f)
            fp64, fp64_src = float(j["fp64_tflops"]), j.get("how", "measured (profiles/fp64_peaks.json)")
    except Exception:
        pass
    return hbm, hbm_src, fp64, fp64_src


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()
        self.period = 0.001            # seconds between samples; raised for the host-timed legs (NVML queries contend for driver locks)

    def _nvml(self):
        """NVML in-process (nvidia_ml_py): a few microseconds per sample.  Spawning nvidia-smi every 100 ms instead re-initialises
        NVML over all GPUs of the box each time and takes driver locks that stall the host-issue-bound e2e pipeline (measured:
        0.57 ms per e2e step without it, 0.7-1.1 ms with it)."""
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self._halt.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            self.rows.append([str(sm), str(mx), str(pw)] + ["Active" if r & b else "Not Active" for _, b in bits])
            self._halt.wait(self.period)

    def run(self):
        try:
            idx = os.environ.get("CUDA_VISIBLE_DEVICES")
            if idx:                                            # NVML enumerates physical devices
                self.index = int(idx.split(",")[self.index])
            return self._nvml()
        except Exception:
            pass
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.1)

    @staticmethod
    def _stats(rows):
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(rows)}

    def finish(self, head: int | None = None):
        """Clocks of the timed region of `value` (the first `head` samples) with the whole run's beside them."""
        self._halt.set()
        self.join(timeout=6)
        out = self._stats(self.rows[:head] if head else self.rows)
        if head:
            out["whole_run"] = self._stats(self.rows)
        return out


def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index` (what `numactl --cpunodebind` per rank does): the pinned
    staging buffers of the e2e leg are then first-touched on the GPU's own NUMA node.  Returns (cpus bound to, all cpus allowed)."""
    allowed = os.sched_getaffinity(0)
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis else index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        words = (max(allowed) + 64) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        local = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1} & allowed
        if local and local != allowed:
            os.sched_setaffinity(0, local)
            return sorted(local), sorted(allowed)
    except Exception:
        pass
    return sorted(allowed), sorted(allowed)


def cpu_baseline(seconds: float = 12.0, threads: int | None = None):
    """The oracle port of the reference solver on the host cores, bounded sample of the same workload."""
    from hierarchicalkarting_b200 import scenarios as S
    from oracle import oracle as O
    threads = threads or os.cpu_count() or 1
    sample = BATCH
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(S.OVAL, sample, 2, seed=20260001))
    O.lqng_solve_batch(A[:256], B[:256], Q[:256], q[:256], R[:256], x0[:256], HORIZON, threads=threads, full=False)
    done, t0 = 0, time.perf_counter()
    while True:
        O.lqng_solve_batch(A, B, Q, q, R, x0, HORIZON, threads=threads, full=False)
        done += sample
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    t1 = time.perf_counter()
    O.lqng_solve_batch(A[:4096], B[:4096], Q[:4096], q[:4096], R[:4096], x0[:4096], HORIZON, threads=1, full=False)
    single = 4096 / (time.perf_counter() - t1)
    return {"value": done / el, "unit": "solves/s", "cores": threads, "kind": "port",
            "sample": f"{done} solves of BASELINE config 2 (the whole 65,536-problem batch, repeated for {el:.1f} s); "
                      f"C oracle restatement of KartLQR.solveFeedbackLQR, OpenMP static split; single thread: {single:.0f} solves/s",
            "single_thread_value": single}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path. The Unity C# sources cannot be compiled here
    (no .NET SDK in the image or on the GPU box), so this arm is the C oracle port with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hierarchicalkarting_b200 import scenarios as S
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    sample = args.batch                                  # the b200 arm's batch: same config, same problems (seed 20260001, rank 0's shard)
    A, B, Q, q, R, x0 = S.assemble_dense(S.make_problems(S.OVAL, sample, 2, seed=20260001))
    for _ in range(max(min(args.warmup, 2), 1)):
        O.lqng_solve_batch(A, B, Q, q, R, x0, HORIZON, threads=threads, full=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.lqng_solve_batch(A, B, Q, q, R, x0, HORIZON, threads=threads, full=False)
    el = time.perf_counter() - t0
    v = sample * args.steps / el
    line = {"impl": "reference", "metric": "lqng_solves_per_s", "value": v, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: 65,536 independent 2-kart LQNG problems per GPU (random linearisation points along "
                                   "the Oval track, seed 20260001+rank), horizon 3, dt=(double)0.02f; output u0 of every player + status",
                       "batch_per_gpu": sample, "horizon": HORIZON, "players": 2,
                       "note": "the CPU arm solves rank 0's whole batch per step on the host cores (N > 1: rank 0 only)"},
            "cpu_baseline": {"value": v, "unit": "solves/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} problems per step x {args.steps} steps; C oracle port of KartLQR.solveFeedbackLQR "
                                       "(reference C# not compilable: no .NET in image), OpenMP over all host threads"},
            "e2e": {"value": v, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_STDOUT_FD = None


def _only_json_on_stdout():
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; emit() writes the one JSON line."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    os.write(_STDOUT_FD if _STDOUT_FD is not None else 1, data)


def main():
    _only_json_on_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mcts", action="store_true")
    ap.add_argument("--no-race", action="store_true")
    ap.add_argument("--no-lqng4", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=2.0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from hierarchicalkarting_b200 import abi, scenarios as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bound_cpus, all_cpus = bind_to_gpu_numa(local)                          # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = abi.load_library()                       # raises if the CUDA library is missing: there is no fallback
    abi.check(lib.hk_init(local))
    dev = torch.device("cuda", local)
    batch, N, n, m = args.batch, 2, 8, 4

    # ---- synthetic shard of this rank (seeded, disjoint per rank) -------------------------------------------------------
    prob = S.make_problems(S.OVAL, batch, N, seed=20260001 + rank)
    host = S.assemble_dense(prob)                                           # A,B,Q,q,R,x0 (numpy, record layout)
    pinned = [torch.from_numpy(a).pin_memory() for a in host]
    h2d_bytes = int(sum(a.nbytes for a in host))
    sets = []
    for k in range(N_INPUT_SETS):                                           # distinct device copies, rotated per step (L2 hygiene)
        sets.append([p.to(dev, non_blocking=True).clone() for p in pinned])
    u0_d = torch.empty((batch, m), dtype=torch.float64, device=dev)
    st_d = torch.empty((batch,), dtype=torch.int32, device=dev)
    u0_h = torch.empty((batch, m), dtype=torch.float64).pin_memory()
    st_h = torch.empty((batch,), dtype=torch.int32).pin_memory()
    d2h_bytes = int(u0_h.numel() * 8 + st_h.numel() * 4)
    # a real (non-NULL) stream: the ABI treats a NULL stream as "the calling thread's own stream", and torch's events
    # only see the stream they are recorded on, so kernel launches and events must share this one
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    torch.cuda.synchronize()                       # input copies above ran on the default stream
    launches = 0

    def step_device(k):
        nonlocal launches
        a = sets[k % N_INPUT_SETS]
        abi.check(lib.hk_lqng_solve_batch_device(batch, N, HORIZON, 0, *[t.data_ptr() for t in a], u0_d.data_ptr(), None, None, None,
                                                 st_d.data_ptr(), stream.cuda_stream))
        launches += 1

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the FP64 roofline denominator of THIS box, burst (clocks at boost, like the short timed region below) ---------------------
    def probe(seconds):
        tf, mhz, ran = (C.c_double(0.0) for _ in range(3))
        abi.check(lib.hk_probe_fp64_peak(seconds, C.byref(tf), C.byref(mhz), C.byref(ran)))
        return {"tflops": tf.value, "sm_mhz_effective": mhz.value, "seconds": ran.value}
    import ctypes as C
    probe(0.01)
    time.sleep(0.5)                                # let the clocks return to idle boost
    peak_burst = probe(0.004)

    # ---- HBM-resident throughput (`value`) -------------------------------------------------------------------------------
    for k in range(args.warmup):
        step_device(k)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches = 0
    k0 = lib.hk_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        step_device(k)
    e1.record(stream)
    e1.synchronize()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    gpu_launches = int(lib.hk_kernel_launch_count() - k0)      # counted inside the library at every <<<>>> site
    # The step is exactly one launch of the LQNG kernel, so the kernel's average launch duration over the timed region is
    # dev_ms / steps (this rank's own events).  Consecutive launches overlap by design: the kernel is launched with programmatic
    # stream serialization, so the next launch's ramp-up fills the SMs its predecessor's tail leaves idle.  The duration of an
    # ISOLATED launch (synchronised on both sides, no overlap) is reported beside it.
    kern_ms = e0.elapsed_time(e1) / args.steps
    per = []
    for k in range(min(args.steps, 20)):
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream); step_device(k); a1.record(stream); a1.synchronize()
        per.append(a0.elapsed_time(a1))
    isolated_ms = float(np.mean(per))
    value_rows = len(sampler.rows)

    # ---- sustained: the same launch back to back for >= 2 s (clocks settle under FP64 power), then the FP64 stream for 2 s ------------
    sustained = None
    if not args.no_sustained:
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        row0 = len(sampler.rows)
        n_sus, t_begin = 0, time.perf_counter()
        s0.record(stream)
        while time.perf_counter() - t_begin < args.sustained_seconds:
            for k in range(256):
                step_device(n_sus + k)
            n_sus += 256
            stream.synchronize()
        s1.record(stream)
        s1.synchronize()
        sus_ms = s0.elapsed_time(s1)
        rows = sampler.rows[row0:]
        tail = rows[len(rows) // 2:] or rows                                   # second half: settled
        peak_sus = probe(args.sustained_seconds)
        sus_tf = batch * FLOPS_PER_SOLVE * n_sus / (sus_ms * 1e-3) / 1e12
        sustained = {"seconds": sus_ms * 1e-3, "launches": n_sus, "solves_per_s": batch * n_sus / (sus_ms * 1e-3), "achieved": sus_tf,
                     "sm_mhz_median_second_half": float(np.median([float(r[0]) for r in tail])) if tail else None,
                     "power_w_median_second_half": float(np.median([float(r[2]) for r in tail])) if tail else None,
                     "power_w_max": float(max(float(r[2]) for r in rows)) if rows else None,
                     "reasons": sorted({n for r in rows for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]) if v.lower().startswith("active")}),
                     "peak": peak_sus}
        launches = 0
    sampler.period = 0.25              # the legs below are timed on the host and are sensitive to driver-lock contention

    # ---- end to end through the host-pointer C-ABI call (`e2e`) ----------------------------------------------------------
    hp = [p.numpy() for p in pinned]
    u0_np, st_np = u0_h.numpy(), st_h.numpy()

    def step_e2e_dense():
        abi.check(lib.hk_lqng_solve_batch(batch, N, HORIZON, 0, *[abi.dptr(a) for a in hp], abi.dptr(u0_np), None, None, None, abi.iptr(st_np)))

    # The reference's call takes provider objects whose only concrete classes are LinearizedBicycle(dt, initial) and
    # LQRCheckpointReachAvoidCost(target, weights, ...): the compact entry ships exactly those constructor arguments
    # (352 B per 2-kart problem instead of the 1,664 B dense record) and assembles A,B,Q,q,R on the GPU.
    ckeys = ("x0", "target", "tw", "cw", "aw", "otgt", "otw")
    cpinned = [torch.from_numpy(np.ascontiguousarray(prob[k], dtype=np.float64)).pin_memory() for k in ckeys]
    cnp = [t.numpy() for t in cpinned]
    compact_bytes = int(sum(a.nbytes for a in cnp))

    def step_e2e():
        abi.check(lib.hk_lqng_assemble_solve_batch(batch, N, HORIZON, float(prob["dt"]), *[abi.dptr(a) for a in cnp], abi.dptr(u0_np), abi.iptr(st_np)))

    def timed(fn):
        for _ in range(max(args.warmup, 20)):                                  # fresh pinned buffers take more than a few copies to reach the link's rate
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        torch.cuda.synchronize()
        el = max_over_ranks(time.perf_counter() - t0)
        barrier()
        return el

    # the same provider arguments as ONE interleaved 352-byte record per problem (what a C# caller fills as an array of blittable structs)
    from hierarchicalkarting_b200 import lqr as LQ
    rec_pinned = torch.from_numpy(LQ.pack_records(prob)).pin_memory()
    rec_np = rec_pinned.numpy()

    def step_e2e_packed():
        abi.check(lib.hk_lqng_assemble_solve_packed(batch, N, HORIZON, float(prob["dt"]), abi.dptr(rec_np), abi.dptr(u0_np), abi.iptr(st_np)))

    # the link before the host-pointer legs (some boxes of the pool slow down after the first tens of ms of sustained traffic: the probe after
    # the legs then reads 15-40 GB/s although the first leg ran at the full rate)
    def _link_gbs():
        h = torch.empty(compact_bytes // 8, dtype=torch.float64).pin_memory()
        d = torch.empty_like(h, device=dev)
        for _ in range(20):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        return 10 * compact_bytes / (time.perf_counter() - t0) / 1e9
    try:
        link_before = _link_gbs()
    except Exception:
        link_before = None
    if os.environ.get("HK_BENCH_PACKED_FIRST"):                               # measurement aid: order of the two compact legs
        e2e_packed_s = timed(step_e2e_packed)
        u0_packed = u0_np.copy()
        e2e_s = timed(step_e2e)
        packed_equal = bool(np.array_equal(u0_packed, u0_np))
    else:
        e2e_s = timed(step_e2e)
        u0_seven = u0_np.copy()
        e2e_packed_s = timed(step_e2e_packed)
        packed_equal = bool(np.array_equal(u0_seven, u0_np))
    e2e_dense_s = timed(step_e2e_dense)

    # what this box's host link gives (the compact e2e call varies 0.57 - 1.4 ms between boxes of the pool at identical code, the 109 MB
    # dense call does not): one 23 MB pinned H2D copy, and the round trip of a small one
    def _probe():
        big_h = torch.empty(compact_bytes // 8, dtype=torch.float64).pin_memory()
        big_d = torch.empty_like(big_h, device=dev)
        small_h = torch.empty(1024, dtype=torch.float64).pin_memory()
        small_d = torch.empty_like(small_h, device=dev)
        for _ in range(20):
            big_d.copy_(big_h, non_blocking=True)
        torch.cuda.synchronize()
        barrier()                                                          # N > 1: every rank copies at the same time (shared uplinks / host memory)
        t0 = time.perf_counter()
        for _ in range(10):
            big_d.copy_(big_h, non_blocking=True)
        torch.cuda.synchronize()
        gbs = 10 * compact_bytes / (time.perf_counter() - t0) / 1e9
        gbs_slowest = -max_over_ranks(-gbs)                                # the slowest rank's rate under that load
        # the call's pipeline copies on four streams at once; what four concurrent copies of a quarter each reach
        streams4 = [torch.cuda.Stream(device=dev) for _ in range(4)]
        q4 = (compact_bytes // 8) // 4
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            for i, st4 in enumerate(streams4):
                with torch.cuda.stream(st4):
                    big_d[i * q4:(i + 1) * q4].copy_(big_h[i * q4:(i + 1) * q4], non_blocking=True)
        torch.cuda.synchronize()
        gbs4 = 10 * 4 * q4 * 8 / (time.perf_counter() - t0) / 1e9
        t0 = time.perf_counter()
        for _ in range(200):
            small_d.copy_(small_h, non_blocking=True)
            torch.cuda.synchronize()
        us = 1e6 * (time.perf_counter() - t0) / 200
        drv = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            drv = nv.nvmlSystemGetDriverVersion()
            drv = drv.decode() if isinstance(drv, bytes) else drv
        except Exception:
            pass
        return {"h2d_23mb_gbs": gbs, "h2d_23mb_gbs_slowest_rank_concurrent": gbs_slowest, "h2d_23mb_4streams_gbs": gbs4, "h2d_8kb_roundtrip_us": us,
                "driver": drv}
    try:
        box_probe = _probe()
        box_probe["h2d_23mb_gbs_before_the_legs"] = link_before
    except Exception as exc:
        box_probe = {"error": repr(exc)[:200]}

    # ---- MCTS rollouts/s (BASELINE config 4: Complex, 2 karts, 10^6 leaf-parallel rollouts per decision) ------------------
    mcts_obj = None
    if not args.no_mcts:
        from hierarchicalkarting_b200 import mcts as M, tracks
        G = M.Game(tracks.COMPLEX, 2, 2)
        leaf = tracks.root_state(tracks.COMPLEX, 3 + rank, [2, 3], teams=[0, 1], tire_age=2500, times=[0, 80])
        for w in range(3):
            G.rollouts(leaf, MCTS_ROLLOUTS, seed=20260003, rollout_offset=rank * MCTS_ROLLOUTS)
        barrier()
        reps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        plies = 0
        for r in range(reps):
            out = G.rollouts(leaf, MCTS_ROLLOUTS, seed=20260003 + r, rollout_offset=rank * MCTS_ROLLOUTS)
            plies += out["plies"]
        el = max_over_ranks(time.perf_counter() - t0)
        mcts_obj = {"metric": "mcts_rollouts_per_s", "value": world * MCTS_ROLLOUTS * reps / el, "unit": "rollouts/s",
                    "plies_per_s": world * plies / el, "ms_per_decision": 1e3 * el / reps, "rollouts_per_decision": MCTS_ROLLOUTS,
                    "config": "BASELINE config 4: Complex, 2 karts, depth 8, bucket 2, Philox4x32-10; host call incl. result D2H",
                    "gpu_launches": reps}
        # the same decision through the reference's own entry points (KartMCTS.constructSearchTree + getBestStatesSequence), both values
        # of its `parallel` argument: false = the sequential search HierarchicalKartAgent runs (one playout per iteration, device-resident
        # tree, hk_mcts_forest_search); true = processLeaf with 4,096 playouts per child (host tree, GPU rollout batches)
        try:
            root_state = M.DiscreteGameState(G, leaf)
            M.KartMCTS.constructSearchTree(root_state, seed=7, max_iterations=32)                 # warm-up
            t0 = time.perf_counter()
            root = M.KartMCTS.constructSearchTree(root_state, seed=20260003, max_iterations=512)
            el_seq = time.perf_counter() - t0
            seq = M.KartMCTS.getBestStatesSequence(root)
            mcts_obj["tree_search"] = {"ms_per_decision": 1e3 * el_seq, "iterations": 512, "episodes_at_root": int(root.numEpisodes),
                                       "nodes": int(root.childrenAsRoot), "best_sequence_states": len(seq), "parallel": False,
                                       "api": "KartMCTS.constructSearchTree(state, parallel=false) + getBestStatesSequence (Python mirror of the C# API "
                                              "over the C-ABI): the reference's sequential search, one GPU thread, incl. the node-record download and the Python "
                                              "KartMCTSNode graph; native_call_ms = hk_mcts_forest_search for this one tree alone"}
            F1 = M.Forest(G, 1, 1 + 512 * 16)                                                     # the native call alone (what the C# shim waits for)
            F1.search([leaf], 32, 7)
            t0 = time.perf_counter()
            F1.search([leaf], 512, 20260003)
            mcts_obj["tree_search"]["native_call_ms"] = 1e3 * (time.perf_counter() - t0)
            F1.close()
            M.KartMCTS.rollouts_per_leaf = 4096
            M.KartMCTS.constructSearchTree(root_state, T=1e9, seed=7, max_iterations=3, parallel=True)          # warm-up
            t0 = time.perf_counter()
            root = M.KartMCTS.constructSearchTree(root_state, T=1e9, seed=20260003, max_iterations=14, parallel=True)
            seq = M.KartMCTS.getBestStatesSequence(root)
            el_tree = time.perf_counter() - t0
            mcts_obj["tree_search_leaf_parallel"] = {"ms_per_decision": 1e3 * el_tree, "iterations": 14, "episodes_at_root": int(root.numEpisodes),
                                                     "nodes": int(root.childrenAsRoot), "best_sequence_states": len(seq), "parallel": True}
            # many trees at once: what config 5 needs at every planning event (32,768 agents)
            roots_b = M._states_array([leaf] * 32768)
            for its in (64, 512):
                F_b = M.Forest(G, 32768, 1 + its * 16)
                F_b.search(roots_b, its, 1)                                    # warm-up at full size: the playout-record buffer is sized by the first call
                t0 = time.perf_counter()
                F_b.search(roots_b, its, 20260003)
                el_b = time.perf_counter() - t0
                mcts_obj.setdefault("forest_search", []).append({"trees": 32768, "iterations": its, "ms": 1e3 * el_b, "decisions_per_s": 32768 / el_b,
                                                                 "playouts_per_s": 32768 * its / el_b,
                                                                 "note": "host call: 5.8 MB of roots up, 92 MB of best states down to pageable memory"})
                F_b.close()
        except Exception as exc:                         # reported, never fatal to the bench line
            mcts_obj["tree_search"] = {"error": repr(exc)[:200]}
        try:                                             # the rollouts are bound by the SM issue rate, not by DRAM
            with open(os.path.join(ROOT, "profiles", "mcts_issue.json")) as f:
                wi = float(json.load(f)["warp_instr_per_rollout"])
            peak_issue = 148 * 4 * 1.965e9                # one warp-instruction per cycle and SM sub-partition at the max SM clock
            ach = mcts_obj["value"] / world * wi
            mcts_obj["roofline"] = {"bound": "issue", "achieved": ach, "peak": peak_issue, "unit": "warp-instr/s", "frac": ach / peak_issue,
                                    "note": f"{wi:.1f} executed warp-instructions per rollout (ncu, profiles/mcts_issue.json) x rollouts/s per GPU"}
        except Exception:
            pass

    # ---- one solve at a time (BASELINE config 1: what HierarchicalKartAgent does at 50 Hz): latency of hk_lqng_solve_one ------------
    one = [np.ascontiguousarray(a[:1]) for a in hp]
    u1 = np.zeros(4)
    for _ in range(20):
        abi.check(lib.hk_lqng_solve_one(N, HORIZON, *[abi.dptr(a) for a in one], abi.dptr(u1)))
    t0 = time.perf_counter()
    for _ in range(200):
        abi.check(lib.hk_lqng_solve_one(N, HORIZON, *[abi.dptr(a) for a in one], abi.dptr(u1)))
    single_us = 1e6 * (time.perf_counter() - t0) / 200

    # ---- 2-kart LQNG with every output (gains P_t, offsets alpha_t, closed-loop rollout: the north_star's wording; the reference
    #      itself returns only u0, which is what `value` measures) -----------------------------------------------------------
    T = HORIZON + 1
    P_d = torch.empty((batch, T, m, n), dtype=torch.float64, device=dev)
    al_d = torch.empty((batch, T, m), dtype=torch.float64, device=dev)
    tr_d = torch.empty((batch, T + 1, n), dtype=torch.float64, device=dev)

    def step_full(k):
        a = sets[k % N_INPUT_SETS]
        abi.check(lib.hk_lqng_solve_batch_device(batch, N, HORIZON, 0, *[t.data_ptr() for t in a], u0_d.data_ptr(), P_d.data_ptr(),
                                                 al_d.data_ptr(), tr_d.data_ptr(), st_d.data_ptr(), stream.cuda_stream))
    for k in range(args.warmup):
        step_full(k)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record(stream)
    for k in range(args.steps):
        step_full(k)
    g1.record(stream)
    g1.synchronize()
    full_ms = max_over_ranks(g0.elapsed_time(g1)) / args.steps
    out_bytes = (T * (m * n + m) + (T + 1) * n + m) * 8 + 4
    full_obj = {"metric": "lqng_full_output_solves_per_s", "value": world * batch / (full_ms * 1e-3), "unit": "solves/s",
                "ms_per_step": full_ms, "outputs": "u0 of every player, P_t [T][4][8], alpha_t [T][4], trajectory [T+1][8], status",
                "hbm_bytes_per_solve": 1664 + out_bytes, "status_nonzero": int(st_d.sum().item()),
                "kernel": "lqng_mma2p_kernel<16, 1, false, true> (FULL mode of the persistent DMMA kernel)"}
    # time-varying operands (A_t, B_t, Q_t, q_t, R_t per stage: 6.5 KB per problem), u0 out — the TMA-staged whole-horizon mode
    rng_tv = np.random.default_rng(20260005 + rank)
    tvh = [np.ascontiguousarray(np.repeat(a[:, None], T, axis=1) * (1.0 + sc * rng_tv.standard_normal((batch, T) + (1,) * (a.ndim - 1))))
           for a, sc in zip(host[:5], (0.01, 0.05, 0.05, 0.05, 0.05))]
    tvd = [torch.from_numpy(a).to(dev) for a in tvh] + [sets[0][5]]
    del tvh

    def step_tv(with_outputs):
        abi.check(lib.hk_lqng_solve_batch_device(batch, N, HORIZON, 1, *[t.data_ptr() for t in tvd], u0_d.data_ptr(),
                                                 P_d.data_ptr() if with_outputs else None, al_d.data_ptr() if with_outputs else None,
                                                 tr_d.data_ptr() if with_outputs else None, st_d.data_ptr(), stream.cuda_stream))
    tv_ms = {}
    for with_outputs in (False, True):
        for _ in range(args.warmup):
            step_tv(with_outputs)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record(stream)
        for _ in range(args.steps):
            step_tv(with_outputs)
        h1.record(stream)
        h1.synchronize()
        tv_ms[with_outputs] = max_over_ranks(h0.elapsed_time(h1)) / args.steps
    tv_obj = {"metric": "lqng_time_varying_solves_per_s", "value": world * batch / (tv_ms[False] * 1e-3), "unit": "solves/s",
              "ms_per_step": tv_ms[False], "with_all_outputs": {"value": world * batch / (tv_ms[True] * 1e-3), "ms_per_step": tv_ms[True]},
              "hbm_bytes_per_solve": T * 1600 + 64 + 36, "status_nonzero": int(st_d.sum().item()),
              "note": "the same 65,536 problems with every stage's operands perturbed (one input set of 428 MB > 126 MB L2)",
              "kernel": "lqng_mma2p_kernel<12, 1, false, *, true> (TV mode: whole horizon staged by TMA)"}
    del P_d, al_d, tr_d, tvd

    # ---- 4-kart LQNG (BASELINE config 3: 1,048,576 Complex 2v2 problems per GPU, HBM-resident) --------------------------------
    lqng4_obj = None
    if not args.no_lqng4:
        uniq, rep = 65536, 16
        b4 = uniq * rep
        # 1,048,576 DISTINCT problems (16 seeded chunks of 65,536, 10.7 GB of dense operands in HBM): generated and assembled chunk by
        # chunk on the host, so that nothing in the launch can be served by L2 from an earlier copy of the same problem
        shapes4 = None
        d4 = None
        for ch in range(rep):
            h4 = S.assemble_dense(S.make_problems(S.COMPLEX, uniq, 4, seed=20260002 + 1000 * ch + rank))
            if d4 is None:
                d4 = [torch.empty((b4,) + a.shape[1:], dtype=torch.float64, device=dev) for a in h4]
            for t, a in zip(d4, h4):
                t[ch * uniq:(ch + 1) * uniq].copy_(torch.from_numpy(a), non_blocking=False)
        del h4
        u4 = torch.empty((b4, 8), dtype=torch.float64, device=dev)
        s4 = torch.empty((b4,), dtype=torch.int32, device=dev)
        torch.cuda.synchronize()

        def step4():
            abi.check(lib.hk_lqng_solve_batch_device(b4, 4, HORIZON, 0, *[t.data_ptr() for t in d4], u4.data_ptr(), None, None, None,
                                                     s4.data_ptr(), stream.cuda_stream))
        step4()
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps4 = 3
        f0.record(stream)
        for _ in range(reps4):
            step4()
        f1.record(stream)
        f1.synchronize()
        ms4 = max_over_ranks(f0.elapsed_time(f1)) / reps4
        bad4 = int(s4.sum().item())
        fp64_pk = peak_burst["tflops"]
        lqng4_obj = {"metric": "lqng4_solves_per_s", "value": world * b4 / (ms4 * 1e-3), "unit": "solves/s", "ms_per_launch": ms4,
                     "batch_per_gpu": b4, "status_nonzero": bad4,
                     "roofline": {"bound": "tensor", "achieved": b4 * FLOPS_PER_SOLVE_4 / (ms4 * 1e-3) / 1e12, "peak": fp64_pk, "unit": "TFLOP/s",
                                  "frac": b4 * FLOPS_PER_SOLVE_4 / (ms4 * 1e-3) / 1e12 / fp64_pk,
                                  "note": f"dense count {FLOPS_PER_SOLVE_4} flops per 4-kart solve (SURVEY.md 8d); kernel lqng_mma4_kernel"},
                     "config": f"BASELINE config 3: 4-kart 2v2 Complex problems, horizon 3; {b4} DISTINCT seeded problems in HBM "
                               f"({sum(t.numel() for t in d4) * 8 / 1e9:.1f} GB of operands per launch, u0 + status out)"}
        del d4, u4, s4
        torch.cuda.empty_cache()
        # the same game through the host-pointer entry: provider constructor arguments (1,280 B per problem) from pinned memory,
        # A,B,Q,q,R assembled on the GPU (10.5 KB dense records stay in HBM), u0 + status back — PCIe-bound
        p4 = S.make_problems(S.COMPLEX, uniq, 4, seed=20260002 + rank)
        c4 = [torch.from_numpy(np.ascontiguousarray(p4[k], dtype=np.float64)).pin_memory() for k in ckeys]
        c4n = [t.numpy() for t in c4]
        u4h = torch.empty((uniq, 8), dtype=torch.float64).pin_memory().numpy()
        s4h = torch.empty((uniq,), dtype=torch.int32).pin_memory().numpy()

        def step4_e2e():
            abi.check(lib.hk_lqng_assemble_solve_batch(uniq, 4, HORIZON, float(p4["dt"]), *[abi.dptr(a) for a in c4n], abi.dptr(u4h), abi.iptr(s4h)))
        for _ in range(3):
            step4_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            step4_e2e()
        el4 = max_over_ranks(time.perf_counter() - t0)
        lqng4_obj["e2e"] = {"value": world * uniq * 10 / el4, "unit": "solves/s", "ms_per_step": 1e2 * el4, "batch_per_gpu": uniq,
                            "h2d_bytes_per_step": int(sum(a.nbytes for a in c4n)), "d2h_bytes_per_step": int(u4h.nbytes + s4h.nbytes),
                            "status_nonzero": int((s4h != 0).sum()),
                            "api": "hk_lqng_assemble_solve_batch, 4 players: pinned provider constructor arguments in, u0 + status out"}
        del c4

    # ---- closed loop without PhysX (BASELINE config 5: 16,384 2-kart Oval races; Fixed high level, LQNG every step) -----------
    race_obj = None
    if not args.no_race:
        from hierarchicalkarting_b200 import race as RC
        RS = RC.Races(S.OVAL, RC.race_params(S.OVAL))
        karts, plans = RC.start_grid(S.OVAL, RACES, seed=20260004 + rank)
        RS.run(karts, plans, 0, 100)                                        # warm-up: the standing start
        barrier()
        k0r = lib.hk_kernel_launch_count()
        t0 = time.perf_counter()
        _, bad = RS.run(karts, plans, 100, RACE_STEPS)
        el = max_over_ranks(time.perf_counter() - t0)
        race_obj = {"metric": "race_agent_steps_per_s", "value": world * 2 * RACES * RACE_STEPS / el, "unit": "agent-steps/s",
                    "races_per_gpu": RACES, "steps": RACE_STEPS, "ms_per_step": 1e3 * el / RACE_STEPS,
                    "realtime_factor": (RACE_STEPS * 0.02) / el * world * RACES,
                    "lqng_status_nonzero": int(bad), "gpu_launches": int(lib.hk_kernel_launch_count() - k0r),
                    "sections_mean": float(karts["section"].mean()),
                    "config": "BASELINE config 5: 2-kart Oval races, kinematic plant, planFixed every 100 steps, one LQNG solve per "
                              "agent and step (recipe + assembly + solve + plant + bookkeeping on the GPU); host call incl. the "
                              "upload and download of the race states; device_resident: the same steps through hk_race_run_device on "
                              "states that stay in HBM; lap / section figures come from the kinematic stand-in for PhysX and the "
                              "raycast-free branches of the target-heading heuristic"}
        dk_r, dp_r = RC.device_state(karts, plans, device=dev)
        RS.run_device(dk_r, dp_r, 100 + RACE_STEPS, 100)
        barrier()
        t0 = time.perf_counter()
        bad_d = RS.run_device(dk_r, dp_r, 200 + RACE_STEPS, RACE_STEPS)
        el_d = max_over_ranks(time.perf_counter() - t0)
        race_obj["device_resident"] = {"value": world * 2 * RACES * RACE_STEPS / el_d, "unit": "agent-steps/s", "ms_per_step": 1e3 * el_d / RACE_STEPS,
                                       "lqng_status_nonzero": int(bad_d), "api": "hk_race_run_device"}
        del dk_r, dp_r

    # ---- the same loop with the MCTS high level (BASELINE config 5 as written: MCTS waypoints -> batched LQNG -> dynamics rollout):
    #      every 100 steps every agent's tree search runs on the GPU (hk_mcts_search_batch, one thread block per tree) --------------
    race_mcts_obj = None
    if not args.no_race and not args.no_mcts:
        try:
            from hierarchicalkarting_b200 import mcts as M2
            prm_m = RC.race_params(S.OVAL, high_mode_mcts=True)
            RM = RC.Races(S.OVAL, prm_m)
            game_m = M2.Game(S.OVAL, 2, prm_m.velocityBucketSize)
            km, pm = RC.start_grid(S.OVAL, RACES, seed=20260004 + rank)
            RM.run(km, pm, 0, 100)                                              # standing start, no plan yet
            blocks_m = 2
            step_ms = race_obj["ms_per_step"] if race_obj else 0.0
            modes = {}
            # the planner as the reference runs it (sequential search, 512 iterations per replan ~ 0.9 s of its C# simulate(), trees kept
            # for 3 cycles, results landing 45 steps = 0.9 s after the search started), and the leaf-parallel mode of round 1
            for name, kw in (("faithful", dict(mode=0, iterations=512, reuse_cycles=3, apply_delay=45)),
                             ("leaf_parallel", dict(mode=1, iterations=24, rollouts_per_leaf=16, reuse_cycles=0, apply_delay=0))):
                kk, pp = km.copy(), pm.copy()
                pl = RC.Planner(game_m, RACES, seed=20260006 + 1000 * rank, **kw)
                RM.run_planned(kk.copy(), pp.copy(), pl, 100, 101)                              # warm-up at full size (one planning event)
                pl.close()
                pl = RC.Planner(game_m, RACES, seed=20260006 + 1000 * rank, **kw)
                barrier()
                k0m = lib.hk_kernel_launch_count()
                t0 = time.perf_counter()
                _, badm = RM.run_planned(kk, pp, pl, 100, 100 * blocks_m)                       # plans at steps 100, 200
                el_m = max_over_ranks(time.perf_counter() - t0)
                _, _, tree_status = pl.state()
                modes[name] = {"value": world * 2 * RACES * 100 * blocks_m / el_m, "unit": "agent-steps/s", "ms_total": 1e3 * el_m,
                               "ms_per_planning_event_derived": (1e3 * el_m - 100 * blocks_m * step_ms) / blocks_m,
                               "plans_per_s_derived": world * 2 * RACES * blocks_m / max(1e-9, el_m - 1e-3 * 100 * blocks_m * step_ms),
                               "planner": {k: v for k, v in kw.items()}, "lqng_status_nonzero": int(badm),
                               "gpu_launches": int(lib.hk_kernel_launch_count() - k0m), "sections_mean": float(kk["section"].mean()),
                               "waypoints_set": int(((pp["lane"] != 0) | (pp["oppLane"] != 0)).sum()),
                               "trees_out_of_nodes": int((tree_status == 3).sum())}
                pl.close()
                if name == "faithful":                                                         # the same two planning events on HBM-resident states
                    pl = RC.Planner(game_m, RACES, seed=20260006 + 1000 * rank, **kw)
                    dk_m, dp_m = RC.device_state(km, pm, device=dev)
                    RM.run_device(dk_m, dp_m, 100, 101, planner=pl)                                 # warm-up: one planning event at full size
                    pl.close()
                    pl = RC.Planner(game_m, RACES, seed=20260006 + 1000 * rank, **kw)
                    dk_m, dp_m = RC.device_state(km, pm, device=dev)
                    barrier()
                    t0 = time.perf_counter()
                    bad_dm = RM.run_device(dk_m, dp_m, 100, 100 * blocks_m, planner=pl)
                    el_dm = max_over_ranks(time.perf_counter() - t0)
                    modes[name]["device_resident"] = {"value": world * 2 * RACES * 100 * blocks_m / el_dm, "unit": "agent-steps/s", "ms_total": 1e3 * el_dm,
                                                      "lqng_status_nonzero": int(bad_dm), "api": "hk_race_run_device"}
                    pl.close()
                    del dk_m, dp_m
            race_mcts_obj = dict(modes["faithful"])
            race_mcts_obj.update({"metric": "race_agent_steps_per_s", "races_per_gpu": RACES, "steps": 100 * blocks_m, "planning_events": blocks_m,
                                  "leaf_parallel": modes["leaf_parallel"],
                                  "config": "BASELINE config 5 with the MCTS high level, GPU-resident (hk_race_run_planned): 2-kart Oval races, every "
                                            "agent replans every 100 steps by the reference's sequential KartMCTS.constructSearchTree (one GPU thread "
                                            "per tree) + getBestStatesSequence, root states and waypoint hand-off as kernels, LQNG every step; one "
                                            "host call incl. the upload and download of the race states; the planning time is derived from the "
                                            "Fixed-mode step time"})
        except Exception as exc:
            race_mcts_obj = {"error": repr(exc)[:300]}

    # ---- Duos: 4-kart 2v2 races on Complex (BASELINE config 3's game inside the loop): N in 1..4 players per problem after the 8 m filter,
    #      solved in the 4-player frame every 4th step; Fixed high level, then the MCTS planner with team scoring ---------------------------
    race4_obj = None
    if not args.no_race:
        try:
            from hierarchicalkarting_b200 import mcts as M2
            R4 = 8192
            RN = RC.RacesN(S.COMPLEX, RC.race_params(S.COMPLEX), 4)
            k4, p4r, b4r, u4r = RC.start_grid_n(S.COMPLEX, R4, 4, seed=20260007 + rank)
            RN.plan_fixed(k4, p4r)
            RN.run_n(k4, p4r, b4r, u4r, 0, 100)                                 # standing start
            barrier()
            k0n = lib.hk_kernel_launch_count()
            t0 = time.perf_counter()
            bad4r = RN.run_n(k4, p4r, b4r, u4r, 100, RACE_STEPS)
            el4r = max_over_ranks(time.perf_counter() - t0)
            npl = RN.recipe_n(k4, p4r, b4r)["n_players"]
            race4_obj = {"metric": "race_agent_steps_per_s", "value": world * 4 * R4 * RACE_STEPS / el4r, "unit": "agent-steps/s", "races_per_gpu": R4,
                         "karts_per_race": 4, "steps": RACE_STEPS, "ms_per_step": 1e3 * el4r / RACE_STEPS, "lqr_every": RN.lqr_every,
                         "lqng_solves_per_s": world * 4 * R4 * (RACE_STEPS // RN.lqr_every) / el4r, "lqng_status_nonzero": int(bad4r),
                         "gpu_launches": int(lib.hk_kernel_launch_count() - k0n), "sections_mean": float(k4["section"].mean()),
                         "players_per_problem_at_end": {str(n): int((npl == n).sum()) for n in (1, 2, 3, 4)},
                         "config": "Duos: 4-kart 2v2 races on Complex (teams [0,0,1,1], start lanes {2,3,2,3} at sections {0,0,1,1}), kinematic plant, "
                                   "planFixed every 100 steps, every agent's LQNG problem (8 m nearby filter, N in 1..4: games of 3-4 in the 4-player frame by "
                                   "lqng_mma4_kernel, games of 1-2 repacked for the 2-kart kernel) every 4th step; host call incl. upload and download of the race states"}
            dk4, dp4, db4 = RC.device_state(k4, p4r, b4r, device=dev)
            du4 = torch.zeros((4 * R4, 8), dtype=torch.float64, device=dev)
            du4[:, :2] = torch.from_numpy(u4r.reshape(-1, 2)).to(dev)
            RN.run_n_device(dk4, dp4, db4, du4, 100 + RACE_STEPS, 100)
            barrier()
            t0 = time.perf_counter()
            bad4d = RN.run_n_device(dk4, dp4, db4, du4, 200 + RACE_STEPS, RACE_STEPS)
            el4d = max_over_ranks(time.perf_counter() - t0)
            race4_obj["device_resident"] = {"value": world * 4 * R4 * RACE_STEPS / el4d, "unit": "agent-steps/s", "ms_per_step": 1e3 * el4d / RACE_STEPS,
                                            "lqng_status_nonzero": int(bad4d), "api": "hk_raceN_run_device"}
            del dk4, dp4, db4, du4
            prm4m = RC.race_params(S.COMPLEX, high_mode_mcts=True)
            RNm = RC.RacesN(S.COMPLEX, prm4m, 4)
            game4 = M2.Game(S.COMPLEX, 4, prm4m.velocityBucketSize)
            km4, pm4, bm4, um4 = RC.start_grid_n(S.COMPLEX, R4, 4, seed=20260007 + rank)
            RNm.run_n(km4, pm4, bm4, um4, 0, 100)
            pl4 = RNm.planner(game4, R4, 256, 20260008 + 1000 * rank, mode=0, reuse_cycles=3, apply_delay=45)
            RNm.run_n(km4.copy(), pm4.copy(), bm4.copy(), um4.copy(), 100, 101, planner=pl4)        # warm-up
            pl4.close()
            pl4 = RNm.planner(game4, R4, 256, 20260008 + 1000 * rank, mode=0, reuse_cycles=3, apply_delay=45)
            barrier()
            t0 = time.perf_counter()
            bad4m = RNm.run_n(km4, pm4, bm4, um4, 100, 200, planner=pl4)
            el4m = max_over_ranks(time.perf_counter() - t0)
            race4_obj["mcts"] = {"value": world * 4 * R4 * 200 / el4m, "unit": "agent-steps/s", "ms_total": 1e3 * el4m, "planning_events": 2,
                                 "ms_per_planning_event_derived": (1e3 * el4m - 200 * race4_obj["ms_per_step"]) / 2,
                                 "planner": {"mode": 0, "iterations": 256, "reuse_cycles": 3, "apply_delay": 45}, "lqng_status_nonzero": int(bad4m),
                                 "waypoints_set": int((pm4["lane"] != 0).sum()), "beliefs_set": int((bm4["lane"] != 0).sum()),
                                 "trees_out_of_nodes": int((pl4.state()[2] == 3).sum())}
            pl4.close()
            km4b, pm4b, bm4b, um4b = RC.start_grid_n(S.COMPLEX, R4, 4, seed=20260007 + rank)
            RNm.run_n(km4b, pm4b, bm4b, um4b, 0, 100)
            def dev4():
                a, b_, c_ = RC.device_state(km4b, pm4b, bm4b, device=dev)
                u_ = torch.zeros((4 * R4, 8), dtype=torch.float64, device=dev)
                u_[:, :2] = torch.from_numpy(um4b.reshape(-1, 2)).to(dev)
                return a, b_, c_, u_
            pl4 = RNm.planner(game4, R4, 256, 20260008 + 1000 * rank, mode=0, reuse_cycles=3, apply_delay=45)
            dk4, dp4, db4, du4 = dev4()
            RNm.run_n_device(dk4, dp4, db4, du4, 100, 101, planner=pl4)                                # warm-up: one planning event at full size
            pl4.close()
            pl4 = RNm.planner(game4, R4, 256, 20260008 + 1000 * rank, mode=0, reuse_cycles=3, apply_delay=45)
            dk4, dp4, db4, du4 = dev4()
            barrier()
            t0 = time.perf_counter()
            bad4dm = RNm.run_n_device(dk4, dp4, db4, du4, 100, 200, planner=pl4)
            el4dm = max_over_ranks(time.perf_counter() - t0)
            race4_obj["mcts"]["device_resident"] = {"value": world * 4 * R4 * 200 / el4dm, "unit": "agent-steps/s", "ms_total": 1e3 * el4dm,
                                                    "lqng_status_nonzero": int(bad4dm), "api": "hk_raceN_run_device"}
            pl4.close()
            del dk4, dp4, db4, du4
        except Exception as exc:
            race4_obj = {"error": repr(exc)[:300]}

    clocks = sampler.finish(value_rows)

    # ---- final gather of per-rank summaries (the only communication) ------------------------------------------------------
    summary = torch.tensor([float(st_d.sum().item()), float(u0_d.sum().item()), float(batch)], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(summary) for _ in range(world)]
        dist.all_gather(gathered, summary)
        summary = torch.stack(gathered).sum(dim=0)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm, hbm_src, fp64_r1, fp64_r1_src = _peaks()
    fp64 = peak_burst["tflops"]                                                # measured on this box, seconds before the timed region
    fp64_src = (f"DMMA m8n8k4 f64 stream on this box (hk_probe_fp64_peak, {1e3 * peak_burst['seconds']:.1f} ms burst at an effective "
                f"{peak_burst['sm_mhz_effective']:.0f} MHz); round-1 pool figure {fp64_r1} TFLOP/s")
    value = world * batch * args.steps / (dev_ms * 1e-3)
    ach_tf = batch * FLOPS_PER_SOLVE / (kern_ms * 1e-3) / 1e12
    ach_gb = batch * (IN_BYTES_PER_SOLVE + OUT_BYTES_PER_SOLVE) / (kern_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "lqng_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": "lqng_solves_per_s", "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "BASELINE config 2: 65,536 independent 2-kart LQNG problems per GPU (random linearisation points along "
                               "the Oval track, seed 20260001+rank), horizon 3, dt=(double)0.02f; output u0 of every player + status",
                   "batch_per_gpu": batch, "horizon": HORIZON, "players": N,
                   "l2": f"{N_INPUT_SETS} distinct device copies of the inputs rotated per step ({N_INPUT_SETS * h2d_bytes / 1e6:.0f} MB > 126 MB L2)",
                   "parallelism": f"independent shards x{world}, no data-path collective"},
        "roofline": {"bound": "tensor", "achieved": ach_tf, "peak": fp64, "unit": "TFLOP/s", "frac": ach_tf / fp64, "traffic": traffic,
                     "note": f"FP64 pipe (DFMA/DMMA share it on B200); algorithmic flops = {FLOPS_PER_SOLVE}/solve (dense count) x {batch} per launch; "
                             f"kernel avg {kern_ms:.4f} ms per launch over the timed region (CUDA events, back-to-back launches overlap their ramp-up "
                             f"with the predecessor's tail through programmatic dependent launch); isolated launch {isolated_ms:.4f} ms; peak = {fp64_src}",
                     "isolated_launch_ms": isolated_ms, "isolated_frac": batch * FLOPS_PER_SOLVE / (isolated_ms * 1e-3) / 1e12 / fp64,
                     "peak_burst": peak_burst,
                     "sustained_frac": (sustained["achieved"] / sustained["peak"]["tflops"]) if sustained else None,
                     "sustained": sustained,
                     "hbm": {"achieved": ach_gb, "peak": hbm, "unit": "GB/s", "frac": ach_gb / hbm, "peak_source": hbm_src,
                             "bytes_per_solve": IN_BYTES_PER_SOLVE + OUT_BYTES_PER_SOLVE}},
        "e2e": {"value": world * batch * args.steps / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": compact_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": 1e3 * e2e_s / args.steps, "host_link": box_probe,
                "pcie_floor_ms": (1e3 * compact_bytes / (max(box_probe["h2d_23mb_gbs"], box_probe["h2d_23mb_4streams_gbs"], link_before or 0.0) * 1e9))
                                 if "h2d_23mb_gbs" in box_probe else None,     # the H2D bytes at the best copy rate this box showed; D2H overlaps
                "api": "hk_lqng_assemble_solve_batch: pinned host buffers holding the reference's provider constructor arguments "
                       "(LinearizedBicycle / LQRCheckpointReachAvoidCost), A,B,Q,q,R assembled on the GPU, u0 + status copied back",
                "packed": {"value": world * batch * args.steps / e2e_packed_s, "unit": "solves/s", "ms_per_step": 1e3 * e2e_packed_s / args.steps,
                           "h2d_bytes_per_step": int(rec_np.nbytes), "d2h_bytes_per_step": d2h_bytes, "equal_to_seven_array_call": packed_equal,
                           "api": "hk_lqng_assemble_solve_packed: the same arguments interleaved per problem (one 352-byte record), one H2D copy per "
                                  "chunk, one TMA bulk copy per problem in the solve kernel"},
                "dense": {"value": world * batch * args.steps / e2e_dense_s, "unit": "solves/s", "h2d_bytes_per_step": h2d_bytes,
                          "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1e3 * e2e_dense_s / args.steps,
                          "api": "hk_lqng_solve_batch: dense A,B,Q,q,R,x0 records from pinned host buffers (PCIe-bound)"}},
        "single_solve_latency_us": single_us,        # BASELINE config 1 through hk_lqng_solve_one (pageable host pointers, H2D + kernel + D2H + sync)
        "gpu_launches": gpu_launches, "clocks": clocks,
        "cpu_affinity": {"bound": len(bound_cpus), "allowed": len(all_cpus),
                         "note": "process pinned to the CPUs NVML reports local to its GPU (numactl-style) for every leg but cpu_baseline"},
        "summary": {"status_nonzero": summary[0].item(), "u0_checksum": summary[1].item(), "problems": summary[2].item()},
    }
    if mcts_obj:
        line["mcts"] = mcts_obj
    if race_obj:
        line["race"] = race_obj
    if race_mcts_obj:
        line["race_mcts"] = race_mcts_obj
    if race4_obj:
        line["race4"] = race4_obj
    line["full_outputs"] = full_obj
    line["time_varying"] = tv_obj
    if lqng4_obj:
        line["lqng4"] = lqng4_obj
    if world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)                                      # the CPU baseline gets every core of the box
        line["cpu_baseline"] = cpu_baseline()
        cb = line["cpu_baseline"]["value"]
        line["e2e"]["ratio_vs_cpu_baseline"] = line["e2e"]["value"] / cb       # hk_lqng_assemble_solve_batch (the headline e2e)
        line["e2e"]["dense"]["ratio_vs_cpu_baseline"] = line["e2e"]["dense"]["value"] / cb   # hk_lqng_solve_batch (dense records)
        line["e2e"]["packed"]["ratio_vs_cpu_baseline"] = line["e2e"]["packed"]["value"] / cb
        line["e2e"]["single_call_ratio_vs_cpu_single_thread"] = (1e6 / single_us) / line["cpu_baseline"]["single_thread_value"]   # hk_lqng_solve_one, what the C# shim calls per agent and step
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
