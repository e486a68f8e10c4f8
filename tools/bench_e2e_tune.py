#!/usr/bin/env python
"""e2e tuning: hk_lqng_assemble_solve_batch (pinned host buffers -> u0) for several HK_E2E_CHUNKS, plus raw pinned H2D bandwidth."""
import os, subprocess, sys
CHILD = r"""
import sys, time, torch, numpy as np
sys.path.insert(0, %r)
from hierarchicalkarting_b200 import abi, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
batch = 65536
prob = S.make_problems(S.OVAL, batch, 2, seed=20260001)
keys = ("x0", "target", "tw", "cw", "aw", "otgt", "otw")
pin = [torch.from_numpy(np.ascontiguousarray(prob[k], dtype=np.float64)).pin_memory() for k in keys]
cnp = [t.numpy() for t in pin]
u0 = torch.empty((batch, 4), dtype=torch.float64).pin_memory().numpy(); st = torch.empty(batch, dtype=torch.int32).pin_memory().numpy()
def step(): abi.check(lib.hk_lqng_assemble_solve_batch(batch, 2, 3, float(prob["dt"]), *[abi.dptr(a) for a in cnp], abi.dptr(u0), abi.iptr(st)))
for _ in range(5): step()
best = 1e9
for rep in range(5):
    t0 = time.perf_counter()
    for _ in range(20): step()
    best = min(best, (time.perf_counter() - t0) / 20)
big = torch.empty(23068672 // 8, dtype=torch.float64).pin_memory(); dbig = torch.empty_like(big, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize(); bw = 20 * big.numel() * 8 / (time.perf_counter() - t0) / 1e9
print('%%.4f ms/step  %%.3e solves/s   (single 23 MB pinned H2D: %%.1f GB/s)  u0 checksum %%.9f status!=0: %%d' %% (best * 1e3, batch / best, bw, float(u0.sum()), int((st != 0).sum())))
"""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# arguments: chunks[:copy_streams[:copy_mode]] ...   (HK_E2E_CHUNKS, HK_E2E_COPY_STREAMS, HK_E2E_COPY_MODE of hk_abi.cu)
for spec in (sys.argv[1:] or ["1", "2", "4", "8", "16"]):
    f = spec.split(":"); f += ["2", "0"][len(f) - 1:]
    env = dict(os.environ, HK_E2E_CHUNKS=f[0], HK_E2E_COPY_STREAMS=f[1], HK_E2E_COPY_MODE=f[2])
    r = subprocess.run([sys.executable, "-c", CHILD % root], env=env, capture_output=True, text=True)
    print("chunks %s copy-streams %s mode %s ->" % tuple(f), r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
