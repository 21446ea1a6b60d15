import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egc_b200
from egc_b200 import _lib
lib = egc_b200.load(); dev="cuda:0"; P=_lib.ptr
def run(x, d_bases, d_lin, algo=_lib.GEMM_3XTF32):
    n, f_in = x.shape; bd = d_bases.shape[1]; hab = d_lin.shape[1]
    wb = torch.zeros(f_in, bd, device=dev); wc = torch.zeros(hab, f_in, device=dev)
    outs = [torch.full(s, -7.0, device=dev) for s in ((n, f_in), (f_in, bd), (hab, f_in), (hab,))]
    nbytes = lib.egc_project_bwd_workspace_bytes(n, f_in, bd, hab)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.egc_project_bwd(P(x), P(wb), P(wc), P(d_bases), P(d_lin), n, f_in, bd, hab, *[P(t) for t in outs], algo, P(ws), nbytes, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return outs, ws
n, f_in, bd, hab = 16, 128, 128, 48
for (k0, m0, n0) in [(0,0,0),(0,1,0),(0,0,1),(1,0,0),(0,4,0),(0,0,4),(8,0,0),(3,37,90),(9,100,130)]:
    x = torch.zeros(n, f_in, device=dev); d1 = torch.zeros(n, bd, device=dev); d2 = torch.zeros(n, hab, device=dev)
    x[k0, m0] = 1.0
    if n0 < bd: d1[k0, n0] = 1.0
    else: d2[k0, n0-bd] = 1.0
    outs, ws = run(x, d1, d2)
    full = torch.cat([outs[1], outs[2].t()], 1)     # [f_in, bd+hab]
    nz = torch.nonzero(full)
    print("one-hot (k,m,n)=", (k0,m0,n0), "-> nonzeros:", nz[:6].tolist(), "vals", full[full!=0][:6].tolist(), "min", float(full.min()))
torch.manual_seed(0)
x = torch.randn(n, f_in, device=dev); d1 = torch.randn(n, bd, device=dev); d2 = torch.randn(n, hab, device=dev)
outs, ws = run(x, d1, d2)
ref = x.double().t() @ torch.cat([d1, d2], 1).double()
full = torch.cat([outs[1], outs[2].t()], 1).double()
print("random: |out| max", float(full.abs().max()), "|ref| max", float(ref.abs().max()), "err", float((full-ref).abs().max()))
part = ws[: 128*176*4].view(torch.float32).view(128,176)
print("partial[0] of cta0 abs max", float(part.abs().max()))
