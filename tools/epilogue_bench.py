"""What the fused epilogue buys: a stack of `EGCBlock`s (conv -> BatchNorm -> ReLU -> dropout -> + identity,
/root/reference/experiments/arxiv/norm_models.py:33-40) on the arxiv-shaped graph in eval mode - the whole tail inside the
aggregation kernel - against the same layers followed by the separate torch ops, inference and frozen-statistics
fine-tuning (forward + backward).  CUDA-event times, one JSON line.   usage: python tools/epilogue_bench.py [--layers 3]"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import egc_b200  # noqa: E402
from egc_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=3)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    w = bench.WORKLOADS["arxiv"]
    n, edge_index = bench.synth_graph("arxiv", 0)
    torch.manual_seed(0)
    blocks = torch.nn.ModuleList([egc_b200.EGCBlock(egc_b200.EGConv(128, 128, aggrs=w["aggrs"], num_heads=w["heads"],
                                                                    num_bases=w["bases"], cached=True), 0.5, True)
                                  for _ in range(args.layers)]).to(dev).eval()
    for blk in blocks:
        with torch.no_grad():
            blk.bn.running_mean.normal_(0, 0.2)
            blk.bn.running_var.uniform_(0.5, 2.0)
    gin = bench.device_graph_input(w, n, edge_index, dev)
    x = torch.randn(n, 128, device=dev)
    go = torch.randn(n, 128, device=dev)

    def fused(h):
        for blk in blocks:
            h = blk(h, gin)
        return h

    def unfused(h):
        for blk in blocks:
            y = F.relu(blk.bn(blk.conv(h, gin)))                 # eval mode: running statistics, dropout = identity
            h = y + h
        return h

    res = {"workload": f"{args.layers} x EGCBlock (EGC-M 128->128 + BatchNorm(eval) + ReLU + residual), arxiv-shaped graph",
           "nodes": n, "layers": args.layers}
    with torch.no_grad():
        a, b = fused(x), unfused(x)
    res["max_rel_diff_fused_vs_unfused"] = bench.rel_err(a, b)
    for name, fn in (("fused", fused), ("unfused", unfused)):
        def infer():
            with torch.no_grad():
                fn(x)
        before = _lib.launch_count()
        ms, _ = bench.cuda_timed(infer, args.steps, args.warmup)
        res[f"inference_ms_{name}"] = ms
        res[f"inference_launches_{name}"] = (_lib.launch_count() - before) / (args.steps + args.warmup)
        params = [p for blk in blocks for p in blk.conv.parameters()]

        def train():
            xs = x.detach().requires_grad_(True)
            torch.autograd.grad(fn(xs), [xs] + params, go)
        ms, _ = bench.cuda_timed(train, args.steps, args.warmup)
        res[f"fwd_bwd_ms_{name}"] = ms
    res["inference_speedup"] = res["inference_ms_unfused"] / res["inference_ms_fused"]
    res["fwd_bwd_speedup"] = res["fwd_bwd_ms_unfused"] / res["fwd_bwd_ms_fused"]
    print(json.dumps(res))


if __name__ == "__main__":
    main()
