"""Timing experiments on the projection kernel: EGC_TC_DEBUG bit mask (see TcParams::debug)."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath("%s"))))
import egc_b200
from egc_b200 import _lib
from egc_b200.functional import project, project_backward
n, f_in, bd, hab = int(sys.argv[1]), 128, int(sys.argv[2]), int(sys.argv[3])
dev = "cuda:0"; torch.manual_seed(0)
x = torch.randn(n, f_in, device=dev); wb = torch.randn(f_in, bd, device=dev) * 0.1
wc = torch.randn(hab, f_in, device=dev) * 0.1; bc = torch.randn(hab, device=dev)
d_bases = torch.randn(n, bd, device=dev); d_lin = torch.randn(n, hab, device=dev)
for _ in range(3):
    b, w = project(x, wb, wc, bc, False, _lib.GEMM_3XTF32)
ref = x.double() @ wb.double()
err = float((b.double() - ref).abs().max() / ref.abs().max())
_lib.profile_enable(True)
for _ in range(10):
    project(x, wb, wc, bc, False, _lib.GEMM_3XTF32)
    project_backward(x, wb, wc, d_bases, d_lin, True, False, False, False, _lib.GEMM_3XTF32)
torch.cuda.synchronize()
prof = _lib.profile_collect()
print("debug", os.environ.get("EGC_TC_DEBUG", "0"), "n", n, "bases err %%.2e" %% err, {k: round(t / c, 4) for k, (c, t) in prof.items()})
''' % os.path.join(here, "x.py")
for shape in (("169343", "128", "48"), ("736389", "64", "32")):
    for dbg in ("0", "1", "2", "4", "8", "3", "7"):
        env = dict(os.environ, EGC_TC_DEBUG=dbg)
        subprocess.run([sys.executable, "-c", code, *shape], env=env)
