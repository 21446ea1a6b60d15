"""Runs the projection kernels alone at the arxiv-shaped size (for ncu captures):  python tools/gemm_only.py [n] [reps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egc_b200
from egc_b200 import _lib
from egc_b200.functional import project, project_backward

n = int(sys.argv[1]) if len(sys.argv) > 1 else 169343
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
f_in, bd, hab = 128, 128, 48
dev = "cuda:0"
torch.manual_seed(0)
x = torch.randn(n, f_in, device=dev); wb = torch.randn(f_in, bd, device=dev) * 0.1
wc = torch.randn(hab, f_in, device=dev) * 0.1; bc = torch.randn(hab, device=dev)
d_bases = torch.randn(n, bd, device=dev); d_lin = torch.randn(n, hab, device=dev)
for _ in range(reps):
    project(x, wb, wc, bc, False, _lib.GEMM_3XTF32)
    project_backward(x, wb, wc, d_bases, d_lin, True, True, True, True, _lib.GEMM_3XTF32)
torch.cuda.synchronize()
print("done")
