import os, sys, faulthandler, atexit
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
print("torch ok", flush=True)
import egc_b200
from egc_b200 import _lib
atexit.register(lambda: print("atexit reached", flush=True))
lib = egc_b200.load()
print("lib ok", lib.egc_build_info().decode(), flush=True)
dev = torch.device("cuda:0")
torch.manual_seed(0)
n, e = 2000, 16000
ei = torch.randint(0, n, (2, e))
aggrs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["symnorm", "max", "std"]
conv = egc_b200.EGConv(64, 128, aggrs=aggrs, num_heads=4, num_bases=4).to(dev)
x = torch.randn(n, 64, device=dev, requires_grad=True)
print("before fwd", flush=True)
with torch.no_grad():
    out = conv(x, ei.to(dev))
print("launched", flush=True)
torch.cuda.synchronize()
print("fwd ok", float(out.abs().sum()), flush=True)
out = conv(x, ei.to(dev))
torch.cuda.synchronize()
print("fwd(train) ok", flush=True)
out.sum().backward()
torch.cuda.synchronize()
print("bwd ok", flush=True)
