"""Projection kernels alone, timed with the library's CUDA-event tracing: python tools/exp_gemm.py [arxiv|mag] [reps]
Diagnostics: EGC_TC_MAX_RAW_STAGES caps the A ring depth (bytes in flight), algo 3 = single-pass TF32."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egc_b200
from egc_b200 import _lib
from egc_b200.functional import project, project_backward

shape = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
n, f_in, bd, hab = (169343, 128, 128, 48) if shape == "arxiv" else (736389, 128, 64, 32)
dev = "cuda:0"
torch.manual_seed(0)
x = torch.randn(n, f_in, device=dev); wb = torch.randn(f_in, bd, device=dev) * 0.1
wc = torch.randn(hab, f_in, device=dev) * 0.1; bc = torch.randn(hab, device=dev)
d_bases = torch.randn(n, bd, device=dev); d_lin = torch.randn(n, hab, device=dev)
lib = _lib.load()
for name, algo in (("3xtf32", _lib.GEMM_3XTF32), ("tf32", _lib.GEMM_TF32)):
    for _ in range(3):
        project(x, wb, wc, bc, False, algo)
        project_backward(x, wb, wc, d_bases, d_lin, True, True, True, True, algo)
    torch.cuda.synchronize()
    lib.egc_profile_enable(1)
    for _ in range(reps):
        project(x, wb, wc, bc, False, algo)
        project_backward(x, wb, wc, d_bases, d_lin, True, True, True, True, algo)
    torch.cuda.synchronize()
    prof = _lib.profile_collect()
    lib.egc_profile_enable(0)
    print(shape, name, "raw_stages_cap", os.environ.get("EGC_TC_MAX_RAW_STAGES", "-"),
          {k: round(v[1] / v[0], 4) for k, v in prof.items()}, flush=True)
