import os, sys, torch
sys.path.insert(0, '/root/repo')
import __graft_entry__ as graft, synthdata
nat = graft.load_package()._native
X = synthdata.matrix_torch(torch, 1 << 17, 1024, 4242, 'cuda')
prm = nat.FitParams(256, 256, 1024, 0, 1, 1, 1023, 0, 0)
nat.fit_device(X, prm); torch.cuda.synchronize()
os.environ["IFB_FIT_DBG"] = "1"
nat.fit_device(X, prm); torch.cuda.synchronize()
