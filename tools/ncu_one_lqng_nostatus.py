"""Write-amplification probe: the 2-kart launch with and without the status output (HK_PROBE_STATUS=0 passes NULL)."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
from hierarchicalkarting_b200 import abi, scenarios as S
lib = abi.load_library(); abi.check(lib.hk_init(0))
dev = torch.device('cuda', 0); batch = 65536
host = S.assemble_dense(S.make_problems(S.OVAL, batch, 2, seed=20260001))
d = [torch.from_numpy(a).to(dev) for a in host]
u0 = torch.empty((batch, 4), dtype=torch.float64, device=dev); st = torch.empty(batch, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
with_status = os.environ.get("HK_PROBE_STATUS", "1") != "0"
for k in range(4):
    abi.check(lib.hk_lqng_solve_batch_device(batch, 2, 3, 0, *[t.data_ptr() for t in d], u0.data_ptr(), None, None, None, st.data_ptr() if with_status else None, s.cuda_stream))
torch.cuda.synchronize()
