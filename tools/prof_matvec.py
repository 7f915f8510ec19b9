"""Two resident H_eff applications at the bench shape (chi = 2048, w = 20; random dense MPO tensors) for ncu:
  ncu --set full --import-source on --clock-control none -k regex:zgemm -c 6 -o gpurun_out/r02_matvec_full python tools/prof_matvec.py [chi] [w]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tensornetworks.jl_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
import tnb200
from tnb200 import _lib
from tnb200.api import GMPS, ProjMPS

chi = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
w = int(sys.argv[2]) if len(sys.argv) > 2 else 20
d = 2
rng = np.random.default_rng(0)
cr = lambda *s: np.asfortranarray(rng.standard_normal(s) + 1j * rng.standard_normal(s))
ctx = tnb200.Context(0)
dims = [1, chi, chi, chi, 1]
psi = GMPS(1, d, [cr(dims[i], d, dims[i + 1]) for i in range(4)], 0, ctx=ctx)
psi.center = 2
M = cr(w, d, d, w)
H = GMPS(2, d, [M[:1], M, M, M[..., :1]], ctx=ctx)
env = ProjMPS(psi, H, psi, center=2)
env.setblock(1, cr(chi, w, chi))
env.setblock(4, cr(chi, w, chi))
th = torch.from_numpy(cr(chi, d, d, chi).reshape(-1, order='F').copy()).cuda()
out = torch.empty_like(th)
_lib.check(ctx.lib.tn_env_product_dev(env.h, C.c_void_p(th.data_ptr()), 0, C.c_void_p(out.data_ptr()), 2))
ctx.sync()
print("done")
