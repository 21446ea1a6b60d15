"""Round-2 additions, GPU side: deterministic min/max routing, the advisor's robustness findings (shape validation,
zero-node inputs, valued adjacency in the paper variant, stale REGConv caches) and the stale-step guard."""
import pytest
import torch

import egc_b200
from egc_b200 import _lib
from egc_b200.functional import aggregate_combine_autograd
from oracle import restatement as R
from tests.test_gpu_parity import assert_close, oracle_and_cuda, run_both
from tests.util import random_graph, rel_err, to_adj_csr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


# ------------------------------------------------------------------------------------------------
# deterministic routing of min/max gradients (EGC_BWD_DETERMINISTIC)
# ------------------------------------------------------------------------------------------------
DET_CONFIGS = [  # f_in, f_out, aggrs, heads, bases, valued adjacency
    (128, 128, ["symnorm", "max", "std"], 4, 4, False),       # BASELINE cfg2 shape (128-bit path, G = 32)
    (128, 128, ["max", "min", "std"], 4, 4, False),           # two routed slots, G = 32
    (64, 64, ["max", "min", "mean"], 4, 4, False),            # two routed slots, G = 16
    (64, 64, ["max", "mean"], 4, 4, False),                   # one routed slot, G = 16
    (32, 40, ["max", "min", "sum"], 4, 3, False),             # B*D = 30: scalar path
    (128, 352, ["max"], 8, 4, False),                         # B*D = 176: two passes, no linear stream (no pass 2)
    (24, 32, ["sum", "max", "min"], 4, 2, True),              # per-nnz linear weights (valued SparseTensor)
]


def _grads(c, x, g_gpu, go):
    xc = x.to(DEV).requires_grad_(True)
    out = c(xc, g_gpu)
    return torch.autograd.grad(out, [xc] + list(c.parameters()), go.to(DEV))


@pytest.mark.parametrize("cfg", DET_CONFIGS, ids=lambda c: f"{c[0]}x{c[1]}-{'+'.join(c[2])}-h{c[3]}b{c[4]}{'-valued' if c[5] else ''}")
def test_deterministic_routing_is_bit_reproducible_and_matches_the_oracle(cfg):
    """max / min are discontinuous: a 1e-6 rounding difference in the projected bases can change a winner and move a
    whole gradient entry (seen here: ~1 flip per 4e5 (row, feature) pairs against the fp64 oracle, a 5e-3 'error' in
    grad_x that is no error).  So the inputs are dyadic - x in multiples of 1/4, basis weights in multiples of 1/16 -
    which makes the projection EXACT in fp32, TF32x3 and fp64 alike: winners are then decided by values that are
    bit-identical on both sides, with plenty of exact ties (coarse grid) settled by nnz position."""
    f_in, f_out, aggrs, h, b, valued = cfg
    n = 3000
    ei = random_graph(n, 30000, seed=71, hub=900)             # hub columns > 256 entries: chunk partials + ordered merge
    o, c = oracle_and_cuda(f_in, f_out, aggrs, h, b, seed=9)
    with torch.no_grad():
        wq = torch.round(o.bases_weight * 16) / 16
        o.bases_weight.copy_(wq)
        c.bases_weight.copy_(wq.to(DEV))
    torch.manual_seed(10)
    x, go = torch.round(torch.randn(n, f_in) * 4) / 4, torch.randn(n, f_out)
    if valued:
        val = torch.randint(1, 5, (ei.size(1),)).float() / 2   # dyadic edge weights too
        rowptr, col, v = to_adj_csr(ei, n, val)
        g_cpu = (rowptr, col, v)
        g_gpu = egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), value=v.to(DEV), sparse_sizes=(n, n), is_sorted=True)
    else:
        g_cpu, g_gpu = ei, ei.to(DEV)
    errs = {}
    for det in (False, True):                                 # the atomic path first: same inputs, same bar
        c.deterministic = det
        res = run_both(o, c, x, g_cpu, g_gpu, go)
        errs[det] = {k: (rel_err(a, b64), max(1e-5, 4 * rel_err(b32, b64))) for k, (a, b64, b32) in res.items()}
    bad = {(det, k): v for det, d in errs.items() for k, v in d.items() if not v[0] < v[1]}
    assert not bad, f"(deterministic, tensor): (error, bar) = {bad}"
    c.deterministic = True
    runs = [_grads(c, x, g_gpu, go) for _ in range(5)]        # only the deterministic path is bit-stable
    for r in runs[1:]:
        for a, bb in zip(r, runs[0]):
            assert torch.equal(a, bb)


def test_deterministic_flag_needs_csr2csc_at_the_c_abi():
    """The C entry point refuses the flag without the CSR-position table instead of silently using atomics."""
    n = 200
    ei = random_graph(n, 1000, seed=72)
    g = egc_b200.GraphStructure.from_edge_index(ei.to(DEV), n, False, True)
    g.ensure_csc()
    desc = egc_b200.make_desc(g, 4, 4, 8, ["max"], False)
    lib = egc_b200.load()
    bases, w = torch.randn(n, 32, device=DEV), torch.randn(n, 16, device=DEV)
    out, _, _, saved, saved_arg = egc_b200.aggregate_combine(desc, g, bases, w, None, want_saved=True)
    go, d_w, d_b = torch.randn(n, 32, device=DEV), torch.empty(n, 16, device=DEV), torch.empty(n, 32, device=DEV)
    nbytes = lib.egc_aggregate_bwd_workspace_bytes(desc, g.csc_plan.struct, _lib.BWD_DETERMINISTIC)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    P = _lib.ptr
    rc = lib.egc_aggregate_bwd(desc, P(g.rowptr), P(g.col), None, P(g.colptr), P(g.rowidx), None, None, None,
                               g.csc_plan.struct, P(bases), P(w), P(saved), P(saved_arg), P(go), None, None, P(d_w), P(d_b), None, None, None,
                               _lib.BWD_DETERMINISTIC, 0, P(ws), nbytes, torch.cuda.current_stream().cuda_stream)
    assert rc != 0 and b"csr2csc" in lib.egc_last_error_string()


# ------------------------------------------------------------------------------------------------
# advisor findings of round 1
# ------------------------------------------------------------------------------------------------
def test_aggregate_combine_autograd_validates_shapes():
    n = 100
    g = egc_b200.GraphStructure.from_edge_index(random_graph(n, 400, seed=73).to(DEV), n, False, True)
    bases, w = torch.randn(n, 32, device=DEV), torch.randn(n, 4 * 2 * 4, device=DEV)
    aggregate_combine_autograd(bases, w, None, g, 4, 4, ["sum", "max"])                       # fine
    with pytest.raises(ValueError):
        aggregate_combine_autograd(bases[:-1], w, None, g, 4, 4, ["sum", "max"])              # wrong n_src
    with pytest.raises(ValueError):
        aggregate_combine_autograd(bases, w[:-1], None, g, 4, 4, ["sum", "max"])              # wrong n_dst
    with pytest.raises(ValueError):
        aggregate_combine_autograd(bases, w[:, :-4].contiguous(), None, g, 4, 4, ["sum", "max"])   # wrong H*A*B
    with pytest.raises(ValueError):
        aggregate_combine_autograd(bases, w, torch.zeros(31, device=DEV), g, 4, 4, ["sum", "max"])  # wrong bias length


def test_zero_node_inputs():
    c = egc_b200.EGConv(8, 16, aggrs=["symnorm", "max"], num_heads=4).to(DEV)
    x = torch.zeros(0, 8, device=DEV, requires_grad=True)
    out = c(x, torch.zeros(2, 0, dtype=torch.long, device=DEV))
    assert out.shape == (0, 16)
    grads = torch.autograd.grad(out.sum(), list(c.parameters()))
    assert all(float(g.abs().sum()) == 0.0 for g in grads)
    # a heterogeneous batch in which one node type has no nodes
    types = ("a", "b")
    rel = (("a", "to", "b"), ("b", "to", "a"))
    m = egc_b200.REGConv(8, 16, 4, 2, node_types=types, edge_types=rel).to(DEV)
    xa = torch.randn(5, 8, device=DEV, requires_grad=True)
    xb = torch.zeros(0, 8, device=DEV, requires_grad=True)
    adj = {rel[0]: egc_b200.SparseTensor(rowptr=torch.zeros(1, dtype=torch.long, device=DEV), col=torch.zeros(0, dtype=torch.long, device=DEV),
                                         sparse_sizes=(0, 5), is_sorted=True),
           rel[1]: egc_b200.SparseTensor(rowptr=torch.zeros(6, dtype=torch.long, device=DEV), col=torch.zeros(0, dtype=torch.long, device=DEV),
                                         sparse_sizes=(5, 0), is_sorted=True)}
    out = m({"a": xa, "b": xb}, adj)
    assert out["a"].shape == (5, 16) and out["b"].shape == (0, 16)
    torch.autograd.grad(out["a"].sum() + out["b"].sum(), [xa] + list(m.parameters()), allow_unused=True)


def test_paper_variant_keeps_adjacency_values_for_the_non_symadd_aggregators():
    """symadd mixed with add / max on a VALUED adj_t and add_self_loops=False: the non-symadd aggregators multiply by the
    adjacency values (ref layers.py:225 `matmul(adj_t, x, reduce=aggr)`), symadd uses gcn_norm of the same values."""
    n = 300
    ei = random_graph(n, 2000, seed=74, self_loops=0, dups=0)
    key = torch.unique(ei[1] * n + ei[0])
    ei = torch.stack([key % n, key // n])
    val = torch.rand(ei.size(1)) + 0.5
    rowptr, col, v = to_adj_csr(ei, n, val)
    torch.manual_seed(3)
    conv = egc_b200.EfficientGraphConv(16, 32, 4, 2, False, add_self_loops=False, aggrs=["symadd", "add", "max"]).to(DEV)
    x = torch.randn(n, 16)
    adj = egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), value=v.to(DEV), sparse_sizes=(n, n), is_sorted=True)
    out = conv(x.to(DEV), adj)
    sd = {k: t.detach().cpu().double() for k, t in conv.state_dict().items()}
    ref = R.paper_forward(x.double(), (rowptr, col, v.double()), [sd[f"bases_weight.{i}"] for i in range(2)],
                          sd["comb_weights.weight"], sd["comb_weights.bias"], sd["bias"], ["symadd", "add", "max"], 4,
                          add_self_loops=False)
    assert rel_err(out, ref) < 1e-5


def test_regconv_rebuilds_a_different_adjacency_under_the_same_key():
    types, rel = ("a",), (("a", "to", "a"),)
    m = egc_b200.REGConv(8, 16, 4, 2, node_types=types, edge_types=rel).to(DEV)
    n = 50
    x = {"a": torch.randn(n, 8, device=DEV)}

    def adj(seed):
        rowptr, col, _ = to_adj_csr(random_graph(n, 300, seed=seed), n)
        return egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), sparse_sizes=(n, n), is_sorted=True)

    a1, a2 = adj(1), adj(2)
    y1 = m(x, {rel[0]: a1})["a"]
    y2 = m(x, {rel[0]: a2})["a"]
    assert rel_err(y2, y1) > 1e-3                      # the second adjacency was really used
    fresh = egc_b200.REGConv(8, 16, 4, 2, node_types=types, edge_types=rel).to(DEV)
    fresh.load_state_dict(m.state_dict())
    assert torch.equal(fresh(x, {rel[0]: a2})["a"], y2)
    assert torch.equal(m(x, {rel[0]: a1})["a"], y1)


# ------------------------------------------------------------------------------------------------
# column phases of the backward (row-partitioned callers: halo columns first, then the own columns)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("aggrs,heads,bd,bases_n", [(["symnorm", "std", "mean"], 4, 128, 4), (["symnorm", "max", "std"], 4, 128, 4),
                                                    (["symnorm"], 8, 64, 4), (["sum", "min", "var"], 4, 30, 3), (["max"], 4, 64, 4)])
@pytest.mark.parametrize("det", [False, True])
def test_backward_column_phases_equal_the_single_call(aggrs, heads, bd, bases_n, det):
    from egc_b200.functional import aggregate_backward, aggregate_combine, make_desc
    n = 2500
    g = egc_b200.GraphStructure.from_edge_index(random_graph(n, 25000, seed=75, hub=700).to(DEV), n, "symnorm" in aggrs, True)
    desc = make_desc(g, heads, bases_n, bd // bases_n, aggrs, False)
    torch.manual_seed(6)
    bases = torch.randn(n, bd, device=DEV)
    w = torch.randn(n, heads * len(aggrs) * bases_n, device=DEV)
    go = torch.randn(n, heads * (bd // bases_n), device=DEV)
    _, _, _, saved, saved_arg = aggregate_combine(desc, g, bases, w, None, want_saved=True)
    flags = _lib.BWD_DETERMINISTIC if det else 0
    ref = aggregate_backward(desc, g, bases, w, saved, saved_arg, go, True, flags, want_lin_colsum=True)
    seen = {}
    for split in (0, 1, 1234, n - 1, n):
        got = aggregate_backward(desc, g, bases, w, saved, saved_arg, go, True, flags, want_lin_colsum=True, col_split=split,
                                 between_phases=lambda d: seen.setdefault(split, d[split:].clone()))
        routed = any(a in ("max", "min") for a in aggrs)
        for a, r in zip(got, ref):
            if routed:                                        # routed-first vs routed-last: one more rounding per entry
                assert rel_err(a, r) < 2e-6
            else:
                assert torch.equal(a, r)
        # what the HEAD phase hands to the exchange (columns >= split) is already final
        if routed and not det:
            assert rel_err(seen[split], ref[1][split:]) < 2e-6 if split < n else True
        else:
            assert torch.equal(seen[split], got[1][split:])


# ------------------------------------------------------------------------------------------------
# fused ReLU epilogue
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(128, 128, ["symnorm", "max", "std"], 4, 4), (128, 128, ["symnorm"], 8, 4),
                                 (64, 84, ["sum", "mean", "min", "var"], 4, 4), (32, 40, ["max", "std", "sum"], 4, 3),
                                 (128, 352, ["mean"], 8, 4)], ids=lambda c: f"{c[0]}x{c[1]}-{'+'.join(c[2])}")
def test_fused_relu_equals_relu_after_the_layer(cfg):
    """conv(x, g, relu=True) == torch.relu(conv(x, g)): identical bits forward, and - because the backward mask
    grad * (out > 0) is what torch applies - identical gradients up to the atomic order of min/max routing."""
    f_in, f_out, aggrs, h, b = cfg
    n = 3000
    ei = random_graph(n, 30000, seed=81, hub=600).to(DEV)
    torch.manual_seed(12)
    c = egc_b200.EGConv(f_in, f_out, aggrs=aggrs, num_heads=h, num_bases=b).to(DEV)
    with torch.no_grad():
        c.bias.uniform_(-0.5, 0.5)
    c.deterministic = True                                    # bit-stable routing: the two runs can be compared exactly
    x, go = torch.randn(n, f_in, device=DEV), torch.randn(n, f_out, device=DEV)
    xa = x.clone().requires_grad_(True)
    ya = torch.relu(c(xa, ei))
    ga = torch.autograd.grad(ya, [xa] + list(c.parameters()), go)
    xb = x.clone().requires_grad_(True)
    yb = c(xb, ei, relu=True)
    gb = torch.autograd.grad(yb, [xb] + list(c.parameters()), go)
    assert torch.equal(ya, yb) and float((yb == 0).float().mean()) > 0.2
    for u, v in zip(ga, gb):
        assert torch.equal(u, v)
    with torch.no_grad():
        assert torch.equal(c(x, ei, relu=True), torch.relu(c(x, ei)))


# ------------------------------------------------------------------------------------------------
# pass 1 / column pass as separate entry points (the T-exchange backward of row-partitioned layers)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cfg", [(["symnorm"], 8, 4, 16), (["sum", "mean"], 4, 4, 32), (["symnorm", "std"], 4, 4, 32),
                                 (["mean", "var"], 4, 3, 10)], ids=lambda c: "+".join(c[0]) + f"-h{c[1]}b{c[2]}d{c[3]}")
def test_pass1_only_plus_column_pass_equals_the_one_call_backward(cfg):
    """egc_aggregate_bwd(EGC_BWD_PASS1_ONLY, tstreams_out) followed by egc_aggregate_bwd_cols over the TRANSPOSED adjacency
    (a CSR whose rows are the source columns) gives the d_bases of the single call (same entries per column; agreement to
    fp32 rounding is what is asserted); so does a split into two entry subsets with EGC_BWD_ACCUMULATE."""
    from egc_b200.dist import n_target_streams, transpose_csr
    from egc_b200.functional import aggregate_backward, aggregate_backward_cols, aggregate_combine
    aggrs, h, b, d = cfg
    n = 3000
    ei = random_graph(n, 30000, seed=95, hub=700).to(DEV)
    g = egc_b200.GraphStructure.from_edge_index(ei, n, True, True)
    desc = egc_b200.make_desc(g, h, b, d, aggrs, False)
    torch.manual_seed(16)
    bases, w = torch.randn(n, b * d, device=DEV), torch.randn(n, h * len(aggrs) * b, device=DEV)
    go = torch.randn(n, h * d, device=DEV)
    _, _, _, saved, saved_arg = aggregate_combine(desc, g, bases, w, None, want_saved=True)
    ref = aggregate_backward(desc, g, bases, w, saved, saved_arg, go, True, want_lin_colsum=True)
    L = n_target_streams(aggrs)
    pad = 5                                                     # extra table rows (a partitioned caller's halo rows)
    t_ext = torch.full((n + pad, L * b * d), float("nan"), device=DEV)
    got = aggregate_backward(desc, g, bases, w, saved, saved_arg, go, True, want_lin_colsum=True, tstreams_out=t_ext)
    assert got[1] is None and torch.equal(got[0], ref[0]) and torch.equal(got[2], ref[2]) and torch.equal(got[3], ref[3])
    assert bool(torch.isfinite(t_ext[:n]).all()) and bool(torch.isnan(t_ext[n:]).all())
    rowptr_t, col_t, sym_t = transpose_csr(g.rowptr.cpu(), g.col.cpu(), g.val_sym.cpu())
    gt = egc_b200.GraphStructure.from_prepared(rowptr_t, col_t, n + pad, val_sym=sym_t, device=DEV)
    desc_t = egc_b200.make_desc(gt, h, b, d, aggrs, False)
    desc_t.n_dst, desc_t.n_src = gt.n_src, gt.n_dst
    d_bases = aggregate_backward_cols(desc_t, gt, t_ext, bases)
    print("column pass bit-equal to the one-call backward:", torch.equal(d_bases, ref[1]))
    assert rel_err(d_bases, ref[1]) < 2e-6
    # two entry subsets (targets below / above n // 2), the second launch accumulating
    keep = col_t < n // 2
    rows = torch.repeat_interleave(torch.arange(n), rowptr_t[1:] - rowptr_t[:-1])
    halves = []
    for m in (keep, ~keep):
        rp = torch.zeros(n + 1, dtype=torch.long)
        rp[1:] = torch.cumsum(torch.bincount(rows[m], minlength=n), 0)
        halves.append(egc_b200.GraphStructure.from_prepared(rp, col_t[m], n + pad, val_sym=sym_t[m], device=DEV))
    two = aggregate_backward_cols(desc_t, halves[0], t_ext, bases)
    two = aggregate_backward_cols(desc_t, halves[1], t_ext, bases, d_bases=two, accumulate=True)
    assert rel_err(two, ref[1]) < 2e-6


@pytest.mark.parametrize("cfg", [(["symnorm"], 8, 4, 16), (["sum", "symnorm"], 4, 4, 32), (["sum"], 4, 3, 10)],
                         ids=lambda c: "+".join(c[0]) + f"-h{c[1]}b{c[2]}d{c[3]}")
def test_forward_over_two_entry_subsets_with_agg_init_equals_one_call(cfg):
    """egc_epilogue.agg_init: a launch over the entries with columns < n / 2 (aggregates only), then one over the rest that
    continues those sums, combines and saves = the single launch over every entry (out, saved; fp32 rounding apart)."""
    from egc_b200.functional import aggregate_combine
    aggrs, h, b, d = cfg
    n = 3000
    ei = random_graph(n, 30000, seed=96, hub=700).to(DEV)
    g = egc_b200.GraphStructure.from_edge_index(ei, n, True, True)
    desc = egc_b200.make_desc(g, h, b, d, aggrs, False)
    torch.manual_seed(17)
    bases, w = torch.randn(n, b * d, device=DEV), torch.randn(n, h * len(aggrs) * b, device=DEV)
    bias = torch.randn(h * d, device=DEV)
    out, _, _, saved, _ = aggregate_combine(desc, g, bases, w, bias, want_saved=True)
    rowptr, col, sym = g.rowptr.cpu().long(), g.col.cpu().long(), g.val_sym.cpu()
    rows = torch.repeat_interleave(torch.arange(n), rowptr[1:] - rowptr[:-1])
    halves = []
    for m in (col < n // 2, col >= n // 2):
        rp = torch.zeros(n + 1, dtype=torch.long)
        rp[1:] = torch.cumsum(torch.bincount(rows[m], minlength=n), 0)
        halves.append(egc_b200.GraphStructure.from_prepared(rp, col[m], n, val_sym=sym[m], device=DEV))
    partial = aggregate_combine(desc, halves[0], bases, None, None, want_out=False, want_agg=True)[1]
    out2, _, _, saved2, _ = aggregate_combine(desc, halves[1], bases, w, bias, want_saved=True,
                                              epilogue=(None, None, None, partial))
    assert rel_err(out2, out) < 2e-6 and rel_err(saved2, saved) < 2e-6
    desc_mean = egc_b200.make_desc(g, h, b, d, ["mean"] * len(aggrs), False)
    with pytest.raises(egc_b200.EGCError):                      # a mean divides by the row's entry count: not continuable
        aggregate_combine(desc_mean, halves[1], bases, w, bias, epilogue=(None, None, None, partial))
