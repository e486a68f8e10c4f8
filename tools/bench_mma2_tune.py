#!/usr/bin/env python
"""Tuning sweep for the 2-kart LQNG throughput kernel: one subprocess per (HK_MMA2_VARIANT, HK_MMA2_MINB, extra env)
combination (the library reads those knobs once), 65,536 Oval problems, 4 rotated input sets, CUDA events.
Usage: python tools/bench_mma2_tune.py [K=V,K=V ...]   each argument is one environment to try."""
import os
import subprocess
import sys

CHILD = r"""
import sys, torch, numpy as np
sys.path.insert(0, %r)
from hierarchicalkarting_b200 import abi, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
dev = torch.device('cuda', 0)
batch = 65536
sets = []
for k in range(4):
    host = S.assemble_dense(S.make_problems(S.OVAL, batch, 2, seed=20260001 + k))
    sets.append([torch.from_numpy(a).to(dev) for a in host])
u0 = torch.empty((batch, 4), dtype=torch.float64, device=dev); st = torch.empty(batch, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s); torch.cuda.synchronize()
def run(k):
    d = sets[k %% 4]
    abi.check(lib.hk_lqng_solve_batch_device(batch, 2, 3, 0, *[t.data_ptr() for t in d], u0.data_ptr(), None, None, None, st.data_ptr(), s.cuda_stream))
for k in range(20): run(k)
best = 1e9
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for k in range(200): run(k)
    e1.record(s); e1.synchronize()
    best = min(best, e0.elapsed_time(e1) / 200)
print('%%.4f ms  %%.4e solves/s  checksum %%.9f bad %%d' %% (best, batch / best * 1e3, float(u0.sum()), int((st != 0).sum())))
"""


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    envs = sys.argv[1:] or ["HK_MMA2_MINB=%d" % m for m in (4, 5, 6, 7, 8)]
    for e in envs:
        env = dict(os.environ)
        for kv in e.split(","):
            if kv:
                k, v = kv.split("=")
                env[k] = v
        r = subprocess.run([sys.executable, "-c", CHILD % root], env=env, capture_output=True, text=True)
        print(e, "->", r.stdout.strip() or r.stderr.strip()[-400:], flush=True)


if __name__ == "__main__":
    main()
