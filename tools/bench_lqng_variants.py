import time, numpy as np, torch, sys
sys.path.insert(0,'/root/repo')
import os
from hierarchicalkarting_b200 import abi, scenarios as S
if os.environ.get("HK_LIB_PATH"): abi.LIB_PATH = os.environ["HK_LIB_PATH"]
lib=abi.load_library(); abi.check(lib.hk_init(0))
dev=torch.device('cuda',0)
ONLY = [int(a) for a in sys.argv[1:]]
for N,track,batch in ((4,S.COMPLEX,65536),(3,S.COMPLEX,65536),(1,S.OVAL,262144),(2,S.OVAL,65536)):
    if ONLY and N not in ONLY: continue
    p=S.make_problems(track,batch,N,seed=1)
    host=S.assemble_dense(p)
    TVM = int(os.environ.get("TV", "0"))
    if TVM:   # time-varying operands: every stage a perturbed copy (symmetric Q kept symmetric)
        rng = np.random.default_rng(3)
        tvf = lambda a, sc: np.ascontiguousarray(np.repeat(a[:, None], 4, axis=1) * (1.0 + sc * rng.standard_normal((batch, 4) + (1,) * (a.ndim - 1))))
        host = [tvf(host[0], 0.01), tvf(host[1], 0.05), tvf(host[2], 0.05), tvf(host[3], 0.05), tvf(host[4], 0.05), host[5]]
    d=[torch.from_numpy(a).to(dev) for a in host]
    u0=torch.empty((batch,2*N),dtype=torch.float64,device=dev); st=torch.empty(batch,dtype=torch.int32,device=dev)
    s=torch.cuda.Stream(); torch.cuda.set_stream(s); torch.cuda.synchronize()
    T=4
    P=torch.empty((batch,T,2*N,4*N),dtype=torch.float64,device=dev); al=torch.empty((batch,T,2*N),dtype=torch.float64,device=dev); tr=torch.empty((batch,T+1,4*N),dtype=torch.float64,device=dev)
    for full in (False,True):
        args=[batch,N,3,TVM]+[t.data_ptr() for t in d]+[u0.data_ptr(), P.data_ptr() if full else None, al.data_ptr() if full else None, tr.data_ptr() if full else None, st.data_ptr(), s.cuda_stream]
        for _ in range(3): abi.check(lib.hk_lqng_solve_batch_device(*args))
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(10): abi.check(lib.hk_lqng_solve_batch_device(*args))
        e1.record(s); e1.synchronize()
        ms=e0.elapsed_time(e1)/10
        print(f"N={N} full={full} batch={batch} {ms:.3f} ms  {batch/ms*1e3:.3e} solves/s", flush=True)
