"""One warm process for `ncu -k regex:lqng_mma4`: 65,536 4-kart Complex problems (u0 only), a few launches."""
import sys, torch
sys.path.insert(0, '.')
from hierarchicalkarting_b200 import abi, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
dev = torch.device('cuda', 0); batch = 65536
host = S.assemble_dense(S.make_problems(S.COMPLEX, batch, 4, seed=20260002))
d = [torch.from_numpy(a).to(dev) for a in host]
u0 = torch.empty((batch, 8), dtype=torch.float64, device=dev); st = torch.empty(batch, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
for k in range(4):
    abi.check(lib.hk_lqng_solve_batch_device(batch, 4, 3, 0, *[t.data_ptr() for t in d], u0.data_ptr(), None, None, None, st.data_ptr(), s.cuda_stream))
torch.cuda.synchronize()
