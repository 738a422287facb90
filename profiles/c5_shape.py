import sys, torch, numpy as np
sys.path.insert(0,'.')
from polars_ds_extension_b200 import device as dev
from polars_ds_extension_b200._lib import lib
n,p=40_000_000,64
Z=torch.randn((p+1,n),device='cuda')
F=dev.to_frame(Z,n=n)
M=torch.empty((p+2,p+2),dtype=torch.float64,device='cuda')
for name,run in (("frame",lambda: dev.moments_frame(F,n,p+1,0,p,p,1,out=M)),("colmajor",lambda: dev.moments(Z[:p],Z[p:],n=n,out=M))):
    for _ in range(2): run()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms=e0.elapsed_time(e1)/5
    print(f"C5-shape p=64 {name}: path={lib().pdsb_last_moments_path()} {ms:.3f} ms for {n} rows -> {n*(p+1)*4/ms/1e6:.0f} GB/s ({n*(p+1)*4/ms/1e6/6570:.2%} of HBM peak), {n/ms*1e3:.3e} rows/s")
