"""Generate tests/golden/*.pt from the UNMODIFIED reference source (run in the build container).

    python -m oracle.make_golden

For each case the reference `EGConv` (/root/reference/experiments/optimized_layers.py, imported
through oracle/shims by oracle/reference_loader.py) is run on seeded inputs in fp32 and fp64;
inputs, parameters, output and every gradient are stored, plus the prepared graph as the
reference's own helpers produce it (gcn_norm / add_remaining_self_loops / fill_diag), so integer
structures can be compared bit-for-bit.  The fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")

# name, N, E, F_in, F_out, H, B, aggrs, add_self_loops, sigmoid, bias, input kind, hub degree
CASES = [
    ("egcs_symnorm_ei",   96,  400, 24,  32, 8, 4, ["symnorm"],                  True,  False, True,  "edge_index", 0),
    ("egcs_symnorm_adj",  96,  400, 24,  32, 8, 4, ["symnorm"],                  True,  False, True,  "adj_t",      0),
    ("egcm_arxiv_ei",    128,  700, 32, 128, 4, 4, ["symnorm", "max", "std"],    True,  False, True,  "edge_index", 0),
    ("egcm_arxiv_adj",   128,  700, 32, 128, 4, 4, ["symnorm", "max", "std"],    True,  False, True,  "adj_t",      0),
    ("all_aggr_ei",       80,  500, 16,  64, 4, 4, ["sum", "mean", "symnorm", "min", "max", "var", "std"], True, False, True, "edge_index", 0),
    ("all_aggr_adj_noloop", 80, 500, 16, 64, 4, 4, ["sum", "mean", "symnorm", "min", "max", "var", "std"], False, False, False, "adj_t", 0),
    ("zinc_d13_sum",      70,  160, 20, 104, 8, 4, ["sum"],                      True,  False, True,  "edge_index", 0),
    ("odd_d21_sigmoid",   64,  300, 12,  84, 4, 4, ["mean", "min", "var"],       True,  True,  True,  "edge_index", 0),
    ("weighted_adj",      72,  360, 16,  32, 4, 2, ["sum", "max", "std"],        True,  False, True,  "adj_t_valued", 0),
    ("hub_rows",         600, 1500, 16,  64, 4, 4, ["symnorm", "max", "std"],    True,  False, True,  "edge_index", 560),
    ("mag_wide_d44",      48,  200, 16, 352, 8, 4, ["mean"],                     True,  False, True,  "adj_t",      0),
]


def make_graph(n, e, hub, gen):
    src = torch.randint(0, n - 2, (e,), generator=gen)       # last two nodes stay isolated
    dst = torch.randint(0, n - 2, (e,), generator=gen)
    src[:5] = dst[:5]                                        # a few pre-existing self-loops
    src = torch.cat([src, src[5:25]])                        # duplicated edges
    dst = torch.cat([dst, dst[5:25]])
    if hub:                                                  # one hub target and one hub source (> chunk size)
        others = torch.randperm(n - 2, generator=gen)[:hub]
        src = torch.cat([src, others, torch.full((hub,), 3)])
        dst = torch.cat([dst, torch.full((hub,), 7), others])
    return torch.stack([src, dst])


def run_case(ref, name, n, e, f_in, f_out, h, b, aggrs, loops, sigmoid, bias, kind, hub):
    gen = torch.Generator().manual_seed(sum(ord(c) for c in name))
    ei = make_graph(n, e, hub, gen)
    x = torch.randn(n, f_in, generator=gen)
    grad_out = torch.randn(n, f_out, generator=gen)
    torch.manual_seed(1234)
    conv = ref.EGConv(f_in, f_out, aggrs=aggrs, num_heads=h, num_bases=b, add_self_loops=loops, bias=bias,
                      sigmoid=sigmoid)
    if bias:
        with torch.no_grad():
            conv.bias.uniform_(-0.5, 0.5)
    value = None
    if kind == "edge_index":
        graph_in = ei
    else:
        perm = (ei[1] * n + ei[0]).argsort(stable=True)      # experiments/utils.py:93
        row, col = ei[0][perm], ei[1][perm]
        if kind == "adj_t_valued":
            value = torch.rand(row.numel(), generator=gen) + 0.5
        graph_in = ref.SparseTensor(row=col, col=row, value=value, sparse_sizes=(n, n), is_sorted=True)

    rec = {"name": name, "n": n, "f_in": f_in, "f_out": f_out, "heads": h, "bases": b, "aggrs": aggrs,
           "add_self_loops": loops, "sigmoid": sigmoid, "kind": kind, "edge_index": ei, "x": x,
           "grad_out": grad_out, "state_dict": {k: v.clone() for k, v in conv.state_dict().items()}}
    if kind != "edge_index":
        rowptr, col_, val_ = graph_in.csr()
        rec["adj_rowptr"], rec["adj_col"], rec["adj_value"] = rowptr.clone(), col_.clone(), val_

    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = ref.EGConv(f_in, f_out, aggrs=aggrs, num_heads=h, num_bases=b, add_self_loops=loops, bias=bias,
                       sigmoid=sigmoid).to(dt)
        m.load_state_dict({k: v.to(dt) for k, v in rec["state_dict"].items()})
        xx = x.to(dt).requires_grad_(True)
        gi = graph_in
        if kind == "adj_t_valued":
            gi = graph_in.set_value(value.to(dt))
        out = m(xx, gi)
        params = list(m.named_parameters())
        grads = torch.autograd.grad(out, [xx] + [p for _, p in params], grad_out.to(dt))
        rec[f"out_{tag}"] = out.detach()
        rec[f"grad_x_{tag}"] = grads[0]
        for (pn, _), g in zip(params, grads[1:]):
            rec[f"grad_{pn}_{tag}"] = g

    # the prepared graph, from the reference's own helpers
    sh = rl.shims()
    sym = "symnorm" in aggrs
    if kind == "edge_index":
        if sym:
            ei2, w = sh.gcn_norm(ei, None, num_nodes=n, improved=False, add_self_loops=loops)
        elif loops:
            ei2, w = sh.add_remaining_self_loops(ei)[0], None
        else:
            ei2, w = ei, None
        rec["prep_edge_index"], rec["prep_weight"] = ei2, w
    else:
        if sym:
            a2 = sh.gcn_norm(graph_in, None, num_nodes=n, improved=False, add_self_loops=loops)
        elif loops:
            a2 = sh.torch_sparse.fill_diag(graph_in, 1.0)
        else:
            a2 = graph_in
        rp, c2, v2 = a2.csr()
        rec["prep_rowptr"], rec["prep_col"], rec["prep_value"] = rp.clone(), c2.clone(), v2
    return rec


# paper variant (experiments/layers.py EfficientGraphConv): name, N, E, F_in, F_out, H, B, aggrs, add_self_loops,
# weight post-processing, bias, input kind
PAPER_CASES = [
    ("paper_softmax_mixed", 120, 600, 24, 128, 4, 4, ["symadd", "max", "mean"], True,  "softmax",  True,  "edge_index"),
    ("paper_sigmoid_std",   100, 500, 16,  64, 4, 4, ["symadd", "std", "max"],  True,  "sigmoid",  True,  "edge_index"),
    ("paper_hardtanh_mol",   90, 400, 20,  64, 8, 4, ["add", "max", "mean"],    True,  "hardtanh", False, "edge_index"),
    ("paper_symadd_adj",     96, 400, 24,  32, 8, 4, ["symadd"],                True,  "none",     True,  "adj_t"),
    ("paper_code_noloop",    80, 450, 16,  64, 4, 4, ["symadd", "max", "min"],  False, "none",     True,  "adj_t"),
]


def run_paper_case(ref, name, n, e, f_in, f_out, h, b, aggrs, loops, post, bias, kind):
    gen = torch.Generator().manual_seed(sum(ord(c) for c in name))
    ei = make_graph(n, e, 0, gen)
    x = torch.randn(n, f_in, generator=gen)
    grad_out = torch.randn(n, f_out, generator=gen)
    kw = dict(softmax_weights=post == "softmax", sigmoid_weights=post == "sigmoid", hardtanh_weights=post == "hardtanh",
              add_self_loops=loops, bias=bias, aggrs=aggrs)
    torch.manual_seed(4321)
    conv = ref.EfficientGraphConv(f_in, f_out, h, b, **kw)
    if bias:
        with torch.no_grad():
            conv.bias.uniform_(-0.5, 0.5)
    if kind == "edge_index":
        graph_in = ei
    else:
        perm = (ei[1] * n + ei[0]).argsort(stable=True)
        graph_in = ref.SparseTensor(row=ei[1][perm], col=ei[0][perm], sparse_sizes=(n, n), is_sorted=True)
    rec = {"name": name, "n": n, "f_in": f_in, "f_out": f_out, "heads": h, "bases": b, "aggrs": aggrs,
           "add_self_loops": loops, "post": post, "bias": bias, "kind": kind, "edge_index": ei, "x": x,
           "grad_out": grad_out, "state_dict": {k: v.clone() for k, v in conv.state_dict().items()}}
    if kind != "edge_index":
        rowptr, col_, _ = graph_in.csr()
        rec["adj_rowptr"], rec["adj_col"] = rowptr.clone(), col_.clone()
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = ref.EfficientGraphConv(f_in, f_out, h, b, **kw).to(dt)
        m.load_state_dict({k: v.to(dt) for k, v in rec["state_dict"].items()})
        xx = x.to(dt).requires_grad_(True)
        out = m(xx, graph_in)
        params = list(m.named_parameters())
        grads = torch.autograd.grad(out, [xx] + [p for _, p in params], grad_out.to(dt))
        rec[f"out_{tag}"] = out.detach()
        rec[f"grad_x_{tag}"] = grads[0]
        for (pn, _), g in zip(params, grads[1:]):
            rec[f"grad_{pn}_{tag}"] = g
    return rec


def main():
    ref = rl.load()
    os.makedirs(OUT_DIR, exist_ok=True)
    for case in PAPER_CASES:
        rec = run_paper_case(ref, *case)
        path = os.path.join(OUT_DIR, f"{case[0]}.pt")
        torch.save(rec, path)
        print(f"{case[0]:24s} out|max|={rec['out_f32'].abs().max():.4f}  {os.path.getsize(path) / 1024:.0f} KiB")
    if os.environ.get("EGC_GOLDEN_PAPER_ONLY"):
        return
    for case in CASES:
        rec = run_case(ref, *case)
        path = os.path.join(OUT_DIR, f"{case[0]}.pt")
        torch.save(rec, path)
        print(f"{case[0]:24s} out|max|={rec['out_f32'].abs().max():.4f}  {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
