"""Quick GPU check of the tcgen05 projection kernels against fp64 torch (run under `timeout`)."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egc_b200
from egc_b200 import _lib
from egc_b200.functional import project

def rel(a, b):
    return float((a.double().cpu() - b).abs().max() / b.abs().max())

dev = "cuda:0"
lib = egc_b200.load()
for (n, f_in, bd, hab) in [(128, 128, 128, 48), (1000, 128, 128, 48), (169343, 128, 128, 48), (5000, 104, 52, 32), (777, 64, 64, 32)]:
    torch.manual_seed(0)
    x, wb = torch.randn(n, f_in), torch.randn(f_in, bd) * 0.1
    wc, bc = torch.randn(hab, f_in) * 0.1, torch.randn(hab)
    for algo, name in ((_lib.GEMM_3XTF32, "3xtf32"), (_lib.GEMM_TF32, "tf32"), (_lib.GEMM_FP32_SIMT, "simt")):
        xd, wbd, wcd, bcd = x.to(dev), wb.to(dev), wc.to(dev), bc.to(dev)
        bases, w = project(xd, wbd, wcd, bcd, False, algo)
        torch.cuda.synchronize()
        e1 = rel(bases, x.double() @ wb.double())
        e2 = rel(w, x.double() @ wc.double().t() + bc.double())
        # timing
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): project(xd, wbd, wcd, bcd, False, algo)
        t0.record()
        for _ in range(10): project(xd, wbd, wcd, bcd, False, algo)
        t1.record(); torch.cuda.synchronize()
        # backward d_x
        d_bases, d_lin = torch.randn(n, bd), torch.randn(n, hab)
        outs = [torch.empty(s, device=dev) for s in ((n, f_in), (f_in, bd), (hab, f_in), (hab,))]
        nbytes = lib.egc_project_bwd_workspace_bytes(n, f_in, bd, hab)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        P = _lib.ptr
        ins = [t.to(dev) for t in (x, wb, wc, d_bases, d_lin)]
        _lib.check(lib.egc_project_bwd(*[P(t) for t in ins], n, f_in, bd, hab, *[P(t) for t in outs], algo, P(ws), nbytes,
                                       torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        e3 = rel(outs[0], d_bases.double() @ wb.double().t() + d_lin.double() @ wc.double())
        e4 = rel(outs[1], x.double().t() @ d_bases.double())
        e5 = rel(outs[2], d_lin.double().t() @ x.double())
        e6 = rel(outs[3], d_lin.double().sum(0))
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(10):
            _lib.check(lib.egc_project_bwd(*[P(t) for t in ins], n, f_in, bd, hab, *[P(t) for t in outs], algo, P(ws), nbytes,
                                           torch.cuda.current_stream().cuda_stream))
        b1.record(); torch.cuda.synchronize()
        print(f"n={n:7d} f_in={f_in} bd={bd} hab={hab} {name:7s} bases {e1:.1e} w {e2:.1e} d_x {e3:.1e} dWb {e4:.1e} dWc {e5:.1e} dbc {e6:.1e} fwd {t0.elapsed_time(t1)/10*1e3:7.1f} us bwd {b0.elapsed_time(b1)/10*1e3:7.1f} us", flush=True)
