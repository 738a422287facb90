import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from polars_ds_extension_b200 import device as dev
from polars_ds_extension_b200._lib import lib
torch.manual_seed(0)
n, p, t = 8192, 4, 1
ld = n
Z = torch.randn((p+t, ld), device='cuda')
X, Y = Z[:p], Z[p:]
Zh = np.concatenate([Z.double().cpu().numpy(), np.ones((1, n))])
ref = Zh @ Zh.T
lib().pdsb_set_moments_path(2)
M = dev.moments(X, Y, n=n)
torch.cuda.synchronize()
M = M.cpu().numpy()
np.set_printoptions(precision=4, linewidth=200, suppress=True)
print("tc:\n", M); print("ref:\n", ref)
print("max rel err", np.max(np.abs(M-ref))/np.max(np.abs(ref)))
