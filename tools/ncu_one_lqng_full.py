"""One warm process for `ncu -k regex:lqng_mma2p`: 65,536 2-kart Oval problems with every output (FULL mode), a few launches."""
import sys, torch
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
dev = torch.device('cuda', 0); batch = 65536
host = S.assemble_dense(S.make_problems(S.OVAL, batch, 2, seed=20260001))
d = [torch.from_numpy(a).to(dev) for a in host]
u0 = torch.empty((batch, 4), dtype=torch.float64, device=dev); st = torch.empty(batch, dtype=torch.int32, device=dev)
P = torch.empty((batch, 4, 4, 8), dtype=torch.float64, device=dev); al = torch.empty((batch, 4, 4), dtype=torch.float64, device=dev)
tr = torch.empty((batch, 5, 8), dtype=torch.float64, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for k in range(6):
    abi.check(lib.hk_lqng_solve_batch_device(batch, 2, 3, 0, *[t.data_ptr() for t in d], u0.data_ptr(), P.data_ptr(), al.data_ptr(), tr.data_ptr(), st.data_ptr(), s.cuda_stream))
torch.cuda.synchronize()
