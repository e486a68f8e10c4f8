import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from hierarchicalkarting_b200 import abi, lqr as LQ, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
batch = 65536
prob = S.make_problems(S.OVAL, batch, 2, seed=20260001)
keys = ("x0", "target", "tw", "cw", "aw", "otgt", "otw")
pin = [torch.from_numpy(np.ascontiguousarray(prob[k], dtype=np.float64)).pin_memory() for k in keys]
arr = [t.numpy() for t in pin]
rec = torch.from_numpy(LQ.pack_records(prob)).pin_memory().numpy()
u0 = torch.empty((batch, 4), dtype=torch.float64).pin_memory().numpy(); st = torch.empty(batch, dtype=torch.int32).pin_memory().numpy()
def seven(): abi.check(lib.hk_lqng_assemble_solve_batch(batch, 2, 3, float(prob["dt"]), *[abi.dptr(a) for a in arr], abi.dptr(u0), abi.iptr(st)))
def packed(): abi.check(lib.hk_lqng_assemble_solve_packed(batch, 2, 3, float(prob["dt"]), abi.dptr(rec), abi.dptr(u0), abi.iptr(st)))
for name, fn in (("packed", packed), ("seven", seven), ("packed", packed), ("seven", seven), ("seven", seven), ("packed", packed)):
    for _ in range(5): fn()
    t0 = time.perf_counter()
    for _ in range(50): fn()
    el = (time.perf_counter() - t0) / 50
    print(f"{name}: {el * 1e3:.3f} ms  {batch / el:.3e} solves/s")
