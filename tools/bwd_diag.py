"""Times the backward of one arxiv-shaped EGC-M layer under the EGC_BWD_* tuning / diagnostic flags."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import egc_b200
from egc_b200 import _lib
from bench import synth_graph, WORKLOADS

wl = sys.argv[1] if len(sys.argv) > 1 else "arxiv"
w = WORKLOADS[wl]
dev = torch.device("cuda:0")
n, ei = synth_graph(wl, 0)
torch.manual_seed(0)
conv = egc_b200.EGConv(w["f_in"], w["f_out"], aggrs=w["aggrs"], num_heads=w["heads"], num_bases=w["bases"], cached=True).to(dev)
x = torch.randn(n, w["f_in"], device=dev, requires_grad=True)
go = torch.randn(n, w["f_out"], device=dev)
params = list(conv.parameters())
g_in = ei.to(dev)
for flags in (0, 2, 4):
    conv.bwd_flags = flags
    for _ in range(3):
        out = conv(x, g_in); torch.autograd.grad(out, [x] + params, go)
    _lib.profile_enable(True)
    for _ in range(10):
        out = conv(x, g_in); torch.autograd.grad(out, [x] + params, go)
    torch.cuda.synchronize()
    prof = _lib.profile_collect(); _lib.profile_enable(False)
    print("flags", flags, {k: round(t / c * (c / 10), 4) for k, (c, t) in prof.items() if "bwd" in k or "colsum" in k or "route" in k})
