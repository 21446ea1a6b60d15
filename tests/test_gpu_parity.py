"""Parity of the sm_100a kernels (called through the C ABI) against the CPU oracle and the golden
vectors generated from the reference source.  Bar (BASELINE.json north_star): bit-exact for CSR
construction, degree, self-loops and argmax indices; <= 1e-5 relative error in fp32 for features and
gradients (relative = max|a-b| / max|b|, see tests/util.rel_err)."""
import itertools

import pytest
import torch

import egc_b200
from egc_b200 import _lib
from egc_b200.functional import aggregate_combine, make_desc, project
from oracle import restatement as R
from tests.util import golden_cases, load_golden, random_graph, rel_err, to_adj_csr

pytestmark = pytest.mark.gpu
TOL = 1e-5
DEV = "cuda:0"


def oracle_and_cuda(f_in, f_out, aggrs, h, b, loops=True, sigmoid=False, bias=True, seed=0, **kw):
    torch.manual_seed(seed)
    o = R.EGConvOracle(f_in, f_out, aggrs=aggrs, num_heads=h, num_bases=b, add_self_loops=loops, bias=bias,
                       sigmoid=sigmoid)
    if bias:
        with torch.no_grad():
            o.bias.uniform_(-0.5, 0.5)
    c = egc_b200.EGConv(f_in, f_out, aggrs=aggrs, num_heads=h, num_bases=b, add_self_loops=loops, bias=bias,
                        sigmoid=sigmoid, **kw)
    c.load_state_dict(o.state_dict())
    return o, c.to(DEV)


def _oracle_run(o, x, graph_cpu, grad_out, dtype):
    od = R.EGConvOracle(o.in_channels, o.out_channels, aggrs=o.aggregators, num_heads=o.num_heads,
                        num_bases=o.num_bases, add_self_loops=o.add_self_loops, bias=o.bias is not None,
                        sigmoid=o.sigmoid).to(dtype)
    od.load_state_dict({k: v.to(dtype) for k, v in o.state_dict().items()})
    xo = x.to(dtype).requires_grad_(True)
    if not isinstance(graph_cpu, torch.Tensor) and graph_cpu[2] is not None:
        graph_cpu = (graph_cpu[0], graph_cpu[1], graph_cpu[2].to(dtype))
    out_o = od(xo, graph_cpu)
    po = list(od.named_parameters())
    go = torch.autograd.grad(out_o, [xo] + [p for _, p in po], grad_out.to(dtype))
    return out_o, po, go


def run_both(o, c, x, graph_cpu, graph_gpu, grad_out):
    """fp32 kernels on the GPU vs the oracle on the CPU.  Returns name -> (cuda, oracle fp64, oracle fp32):
    the fp64 run is the truth, the fp32 run shows how much rounding noise the reference's own fp32
    arithmetic carries on this input (std's var = E[x^2] - E[x]^2 cancels catastrophically, so its
    gradient can be off by 1e-4 in fp32 on BOTH sides)."""
    out_o, po, go = _oracle_run(o, x, graph_cpu, grad_out, torch.float64)
    out_s, _, gs = _oracle_run(o, x, graph_cpu, grad_out, torch.float32)
    xc = x.to(DEV).requires_grad_(True)
    out_c = c(xc, graph_gpu)
    pc = dict(c.named_parameters())
    gc = torch.autograd.grad(out_c, [xc] + [pc[n] for n, _ in po], grad_out.to(DEV))
    res = {"out": (out_c, out_o, out_s), "grad_x": (gc[0], go[0], gs[0])}
    for (n, _), a, bb, cc in zip(po, gc[1:], go[1:], gs[1:]):
        res["grad_" + n] = (a, bb, cc)
    return res


def tolerance(ref32, ref64):
    """1e-5 relative, or 4x the reference's own fp32-vs-fp64 rounding error where that is larger."""
    return max(TOL, 4.0 * rel_err(ref32, ref64))


def assert_close(res):
    for k, (a, b64, b32) in res.items():
        e, tol = rel_err(a, b64), tolerance(b32, b64)
        assert e < tol, f"{k}: relative error {e:.3e} >= {tol:.3e}"


# ------------------------------------------------------------------------------------------------
# integer structures: bit-exact
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("symnorm,loops", [(True, True), (True, False), (False, True), (False, False)])
@pytest.mark.parametrize("hub", [0, 700])
def test_csr_from_edge_index_bit_exact(symnorm, loops, hub):
    n = 1000
    ei = random_graph(n, 6000, seed=11, hub=hub, isolated=3)
    g = egc_b200.GraphStructure.from_edge_index(ei.to(DEV), n, symnorm, loops)
    o = R.graph_from_edge_index(ei, n, symnorm, loops)
    assert g.nnz == o.nnz
    assert torch.equal(g.rowptr.cpu().long(), o.rowptr)
    assert torch.equal(g.col.cpu().long(), o.col)
    if symnorm:
        assert torch.equal(g.deg.cpu(), o.deg)
        assert torch.equal(g.dis.cpu(), o.dis)
        assert torch.equal(g.val_sym.cpu(), o.val_sym)
    assert g.max_deg == int((o.rowptr[1:] - o.rowptr[:-1]).max())
    assert (g.plan.n_long > 0) == (g.max_deg > _lib.EGC_CHUNK_EDGES)


@pytest.mark.parametrize("symnorm,loops,valued", [(True, True, False), (True, True, True), (False, True, False),
                                                  (False, True, True), (False, False, True), (True, False, False)])
def test_csr_fill_diag_bit_exact(symnorm, loops, valued):
    n = 800
    ei = random_graph(n, 5000, seed=12, hub=400)
    val = torch.rand(ei.size(1)) + 0.5 if valued else None
    rowptr, col, v = to_adj_csr(ei, n, val)
    g = egc_b200.GraphStructure.from_csr(rowptr.to(DEV), col.to(DEV), v.to(DEV) if valued else None, n, symnorm, loops)
    o = R.graph_from_csr(rowptr, col, v, n, symnorm, loops, True)
    assert torch.equal(g.rowptr.cpu().long(), o.rowptr)
    assert torch.equal(g.col.cpu().long(), o.col)
    if symnorm:
        if valued:      # degree = sum of fp32 values: warp-tree vs sequential order, not bit-exact
            assert rel_err(g.val_sym, o.val_sym) < 1e-6
        else:
            assert torch.equal(g.val_sym.cpu(), o.val_sym)
    elif valued:
        assert torch.equal(g.val_lin.cpu(), o.val_lin)


def test_unsorted_csr_and_bad_ids_raise():
    rowptr = torch.tensor([0, 2, 3], device=DEV)
    with pytest.raises(ValueError, match="sorted"):
        egc_b200.GraphStructure.from_csr(rowptr, torch.tensor([1, 0, 1], device=DEV), None, 2, False, True)
    with pytest.raises(IndexError):
        egc_b200.GraphStructure.from_edge_index(torch.tensor([[0, 5], [1, 0]], device=DEV), 3, True, True)


def test_transpose_and_plan_bit_exact():
    n = 1200
    ei = random_graph(n, 7000, seed=13, hub=900)
    g = egc_b200.GraphStructure.from_edge_index(ei.to(DEV), n, True, True)
    g.ensure_csc()
    o = R.graph_from_edge_index(ei, n, True, True)
    perm = torch.argsort(o.col, stable=True)                       # csr2csc: stable by source
    assert torch.equal(g.csr2csc.cpu().long(), perm)
    assert torch.equal(g.rowidx.cpu().long(), o.row[perm])
    colptr = torch.zeros(n + 1, dtype=torch.long)
    colptr[1:] = torch.cumsum(torch.bincount(o.col, minlength=n), 0)
    assert torch.equal(g.colptr.cpu().long(), colptr)
    assert torch.equal(g.csc_val_sym.cpu(), o.val_sym[perm])
    for plan, ptr_ in ((g.plan, o.rowptr), (g.csc_plan, colptr)):
        deg = ptr_[1:] - ptr_[:-1]
        long_rows = torch.nonzero(deg > _lib.EGC_CHUNK_EDGES).flatten()
        assert plan.n_long == long_rows.numel() and plan.n_long > 0
        assert torch.equal(plan.long_rows.cpu().long(), long_rows)
        chunks = (deg[long_rows] + _lib.EGC_CHUNK_EDGES - 1) // _lib.EGC_CHUNK_EDGES
        assert plan.n_chunks == int(chunks.sum())
        assert torch.equal(plan.long_chunk_ptr.cpu().long(), torch.cat([chunks.new_zeros(1), chunks.cumsum(0)]))
        exp_row = torch.repeat_interleave(long_rows, chunks)
        assert torch.equal(plan.chunk_row.cpu().long(), exp_row)
        k = torch.arange(int(chunks.sum())) - torch.repeat_interleave(chunks.cumsum(0) - chunks, chunks)
        assert torch.equal(plan.chunk_begin.cpu().long(), ptr_[exp_row] + k * _lib.EGC_CHUNK_EDGES)


# ------------------------------------------------------------------------------------------------
# aggregation stage: values and argmax source ids
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bd,dim", [(128, 32), (64, 16), (52, 13), (176, 44), (30, 10), (8, 2)])
@pytest.mark.parametrize("hub", [0, 600])
def test_aggregate_stage_and_argmax_ids(bd, dim, hub):
    n, b = 900, bd // dim
    aggrs = ["sum", "mean", "symnorm", "min", "max", "var", "std"]
    ei = random_graph(n, 8000, seed=21, hub=hub)
    g = egc_b200.GraphStructure.from_edge_index(ei.to(DEV), n, True, True)
    o = R.graph_from_edge_index(ei, n, True, True)
    torch.manual_seed(3)
    bases = torch.randn(n, bd)
    bases[5] = bases[9]                                            # exact ties between two sources
    desc = make_desc(g, 4, b, dim, aggrs, False)
    _, agg, arg, _, _ = aggregate_combine(desc, g, bases.to(DEV), None, None, want_out=False, want_agg=True,
                                            want_arg=True)
    agg_o, arg_o = R.aggregate(o, bases.double(), aggrs)
    assert rel_err(agg, agg_o) < TOL
    for name in ("min", "max"):
        k = aggrs.index(name)
        src_gpu = g.source_ids(arg[:, k]).cpu()
        pos = arg_o[name]
        src_o = torch.where(pos < o.nnz, o.col[pos.clamp(max=o.nnz - 1)], torch.full_like(pos, -1))
        assert torch.equal(src_gpu, src_o), f"argmax source ids differ for {name}"
    if bd == 128 and hub == 0:      # one neighbour per step, sequential fp32 order == the oracle's order
        agg32, _ = R.aggregate(o, bases, aggrs)
        for k, name in enumerate(aggrs):
            if name in ("sum", "symnorm", "min", "max"):
                assert torch.equal(agg[:, k].cpu(), agg32[:, k]), f"{name} not bit-identical to the fp32 oracle"


def test_degree_properties_full_arxiv_size():
    """Size-independent properties at BASELINE cfg2 size: aggregating all-ones gives the degree (sum), 1 (mean,
    max, min), 0 (var), sqrt(1e-5) (std) and sum_j w_ij (symnorm)."""
    from bench import synth_graph
    n, ei = synth_graph("arxiv", seed=0)
    g = egc_b200.GraphStructure.from_edge_index(ei.to(DEV), n, True, True)
    aggrs = ["sum", "mean", "symnorm", "min", "max", "var", "std"]
    desc = make_desc(g, 4, 4, 32, aggrs, False)
    ones = torch.ones(n, 128, device=DEV)
    agg = aggregate_combine(desc, g, ones, None, None, want_out=False, want_agg=True)[1]
    deg = (g.rowptr[1:] - g.rowptr[:-1]).float()
    assert torch.equal(agg[:, 0], deg.view(-1, 1).expand(-1, 128))
    for k in (1, 3, 4):
        assert torch.equal(agg[:, k], torch.ones_like(agg[:, k]))
    assert float(agg[:, 5].abs().max()) == 0.0
    assert torch.allclose(agg[:, 6], torch.full_like(agg[:, 6], 1e-5 ** 0.5))
    rowsum = torch.zeros(n, device=DEV, dtype=torch.float64).index_add_(0, torch.repeat_interleave(
        torch.arange(n, device=DEV), (g.rowptr[1:] - g.rowptr[:-1]).long()), g.val_sym.double())
    assert rel_err(agg[:, 2, 0], rowsum) < TOL
    # linearity of the sum aggregator at full size
    torch.manual_seed(0)
    xa, xb = torch.randn(n, 128, device=DEV), torch.randn(n, 128, device=DEV)
    d1 = make_desc(g, 4, 4, 32, ["sum"], False)
    s = lambda t: aggregate_combine(d1, g, t, None, None, want_out=False, want_agg=True)[1]  # noqa: E731
    assert rel_err(s(2 * xa + xb), 2 * s(xa) + s(xb)) < 1e-5


# ------------------------------------------------------------------------------------------------
# projections
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,f_in,bd,hab", [(1000, 128, 128, 48), (777, 100, 52, 32), (130, 24, 176, 96), (5, 7, 9, 3)])
@pytest.mark.parametrize("sigmoid", [False, True])
def test_project_fwd_bwd(n, f_in, bd, hab, sigmoid):
    torch.manual_seed(0)
    x, wb = torch.randn(n, f_in), torch.randn(f_in, bd) * 0.1
    wc, bc = torch.randn(hab, f_in) * 0.1, torch.randn(hab)
    bases, w = project(x.to(DEV), wb.to(DEV), wc.to(DEV), bc.to(DEV), sigmoid, _lib.GEMM_FP32_SIMT)
    b_o, w_o = R.project(x.double(), wb.double(), wc.double(), bc.double(), sigmoid)
    assert rel_err(bases, b_o) < TOL and rel_err(w, w_o) < TOL
    lib = egc_b200.load()
    d_bases, d_lin = torch.randn(n, bd), torch.randn(n, hab)
    outs = [torch.empty(s, device=DEV) for s in ((n, f_in), (f_in, bd), (hab, f_in), (hab,))]
    nbytes = lib.egc_project_bwd_workspace_bytes(n, f_in, bd, hab)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
    P = _lib.ptr
    dev_in = [t.to(DEV) for t in (x, wb, wc, d_bases, d_lin)]
    _lib.check(lib.egc_project_bwd(*[P(t) for t in dev_in], n, f_in, bd, hab, *[P(t) for t in outs],
                                   _lib.GEMM_FP32_SIMT, P(ws), nbytes, torch.cuda.current_stream().cuda_stream))
    xd, db, dl = x.double(), d_bases.double(), d_lin.double()
    assert rel_err(outs[0], db @ wb.double().t() + dl @ wc.double()) < TOL
    assert rel_err(outs[1], xd.t() @ db) < TOL
    assert rel_err(outs[2], dl.t() @ xd) < TOL
    assert rel_err(outs[3], dl.sum(0)) < TOL


# ------------------------------------------------------------------------------------------------
# whole layer, forward + backward
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_cases())
def test_layer_matches_reference_golden(name):
    rec = load_golden(name)
    c = egc_b200.EGConv(rec["f_in"], rec["f_out"], aggrs=rec["aggrs"], num_heads=rec["heads"],
                        num_bases=rec["bases"], add_self_loops=rec["add_self_loops"],
                        bias="bias" in rec["state_dict"], sigmoid=rec["sigmoid"])
    c.load_state_dict(rec["state_dict"])            # the reference's own state_dict keys / shapes
    c = c.to(DEV)
    if rec["kind"] == "edge_index":
        gi = rec["edge_index"].to(DEV)
    else:
        v = rec["adj_value"]
        gi = egc_b200.SparseTensor(rowptr=rec["adj_rowptr"].to(DEV), col=rec["adj_col"].to(DEV),
                                   value=v.to(DEV) if v is not None else None, sparse_sizes=(rec["n"], rec["n"]),
                                   is_sorted=True)
    x = rec["x"].to(DEV).requires_grad_(True)
    out = c(x, gi)
    names = [n for n, _ in c.named_parameters()]
    grads = torch.autograd.grad(out, [x] + list(c.parameters()), rec["grad_out"].to(DEV))
    assert rel_err(out, rec["out_f64"]) < tolerance(rec["out_f32"], rec["out_f64"])
    assert rel_err(grads[0], rec["grad_x_f64"]) < tolerance(rec["grad_x_f32"], rec["grad_x_f64"])
    for pn, g in zip(names, grads[1:]):
        assert rel_err(g, rec[f"grad_{pn}_f64"]) < tolerance(rec[f"grad_{pn}_f32"], rec[f"grad_{pn}_f64"]), pn


# ------------------------------------------------------------------------------------------------
# paper variant `EfficientGraphConv` (reference experiments/layers.py) through the adapter
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", golden_cases("paper_"))
def test_paper_variant_matches_reference_golden(name):
    rec = load_golden(name)
    post = rec["post"]
    c = egc_b200.EfficientGraphConv(rec["f_in"], rec["f_out"], rec["heads"], rec["bases"], post == "softmax",
                                    add_self_loops=rec["add_self_loops"], bias=rec["bias"], aggrs=rec["aggrs"],
                                    sigmoid_weights=post == "sigmoid", hardtanh_weights=post == "hardtanh")
    c.load_state_dict(rec["state_dict"])            # the reference's own keys: comb_weights.*, bases_weight.<b>, bias
    c = c.to(DEV)
    if rec["kind"] == "edge_index":
        gi = rec["edge_index"].to(DEV)
    else:
        gi = egc_b200.SparseTensor(rowptr=rec["adj_rowptr"].to(DEV), col=rec["adj_col"].to(DEV),
                                   sparse_sizes=(rec["n"], rec["n"]), is_sorted=True)
    x = rec["x"].to(DEV).requires_grad_(True)
    out = c(x, gi)
    names = [n for n, _ in c.named_parameters()]
    grads = torch.autograd.grad(out, [x] + list(c.parameters()), rec["grad_out"].to(DEV))
    assert rel_err(out, rec["out_f64"]) < tolerance(rec["out_f32"], rec["out_f64"])
    assert rel_err(grads[0], rec["grad_x_f64"]) < tolerance(rec["grad_x_f32"], rec["grad_x_f64"])
    for pn, g in zip(names, grads[1:]):
        assert rel_err(g, rec[f"grad_{pn}_f64"]) < tolerance(rec[f"grad_{pn}_f32"], rec[f"grad_{pn}_f64"]), pn


def test_paper_checkpoint_converts_to_egconv():
    """App. B: with add_self_loops=False the two layers coincide under the block / row permutation."""
    n = 400
    ei = random_graph(n, 3000, seed=77)
    aggrs_p, aggrs_e = ["symadd", "max", "mean"], ["symnorm", "max", "mean"]
    torch.manual_seed(5)
    paper = egc_b200.EfficientGraphConv(48, 64, 4, 4, False, add_self_loops=False, aggrs=aggrs_p).to(DEV)
    conv = egc_b200.EGConv(48, 64, aggrs=aggrs_e, num_heads=4, num_bases=4, add_self_loops=False).to(DEV)
    conv.load_state_dict(egc_b200.convert_paper_state_dict(paper.state_dict(), 4, 4, 3))
    x = torch.randn(n, 48, device=DEV)
    assert rel_err(conv(x, ei.to(DEV)), paper(x, ei.to(DEV)).double().cpu()) < TOL
    with pytest.raises(NotImplementedError):        # ref layers.py:222-224
        rowptr, col, _ = to_adj_csr(ei, n)
        egc_b200.EfficientGraphConv(48, 64, 4, 4, False, aggrs=["std"]).to(DEV)(
            x, egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), sparse_sizes=(n, n), is_sorted=True))


CONFIGS = [  # f_in, f_out, aggrs, heads, bases
    (128, 128, ["symnorm", "max", "std"], 4, 4),          # BASELINE cfg2/3 (EGC-M arxiv)
    (128, 128, ["symnorm"], 8, 4),                        # cfg4 (EGC-S mag)
    (104, 104, ["sum"], 8, 4),                            # cfg1 (ZINC, D=13)
    (128, 352, ["mean"], 8, 4),                           # mag output layer, D=44 (two passes)
    (64, 84, ["sum", "mean", "min", "var"], 4, 4),        # D=21
    (32, 40, ["max", "min", "std", "mean", "symnorm", "sum", "var"], 4, 3),   # B*D = 30 (scalar path)
    (48, 48, ["std", "std", "max"], 2, 1),                # duplicates, single basis
]


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: f"{c[0]}x{c[1]}-{'+'.join(c[2])}-h{c[3]}b{c[4]}")
@pytest.mark.parametrize("kind", ["edge_index", "adj_t"])
def test_layer_vs_oracle(cfg, kind):
    f_in, f_out, aggrs, h, b = cfg
    n = 3000
    ei = random_graph(n, 30000, seed=31, hub=800)
    o, c = oracle_and_cuda(f_in, f_out, aggrs, h, b, seed=7)
    torch.manual_seed(8)
    x, go = torch.randn(n, f_in), torch.randn(n, f_out)
    if kind == "edge_index":
        res = run_both(o, c, x, ei, ei.to(DEV), go)
    else:
        rowptr, col, _ = to_adj_csr(ei, n)
        adj = egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), sparse_sizes=(n, n), is_sorted=True)
        res = run_both(o, c, x, (rowptr, col, None), adj, go)
    assert_close(res)


def _degree_pattern_graph(n, degrees, seed, symmetric):
    """Row i gets degrees[i % len(degrees)] distinct sources: exercises the row-/column-block kernels' window
    splits (8 consecutive rows > 384 nnz), rows of exactly 256 / 257 nnz (chunk threshold), runs of empty rows."""
    gen = torch.Generator().manual_seed(seed)
    src, dst = [], []
    for i in range(n):
        d = degrees[i % len(degrees)]
        if d:
            src.append(torch.randperm(n, generator=gen)[:d])
            dst.append(torch.full((d,), i, dtype=torch.long))
    ei = torch.stack([torch.cat(src), torch.cat(dst)])
    if symmetric:                                    # the CSC pass then sees the same column lengths
        ei = torch.cat([ei, ei.flip(0)], 1)
    return ei


BLOCK_CONFIGS = [  # f_in, f_out, aggrs, heads, bases: a specialised shape, a dynamic G = 32 shape, a G = 16 shape
    (128, 128, ["symnorm", "max", "std"], 4, 4),
    (64, 128, ["sum", "mean", "min", "var"], 4, 4),
    (128, 128, ["symnorm"], 8, 4),
    (64, 48, ["max", "mean"], 4, 3),                 # B*D = 36
]


@pytest.mark.parametrize("cfg", BLOCK_CONFIGS, ids=lambda c: f"{c[0]}x{c[1]}-{'+'.join(c[2])}-h{c[3]}b{c[4]}")
@pytest.mark.parametrize("pattern", ["windows", "threshold", "empty_runs"])
@pytest.mark.parametrize("loops", [True, False])
def test_block_kernels_on_degree_patterns(cfg, pattern, loops):
    f_in, f_out, aggrs, h, b = cfg
    degrees = {"windows": [200, 1, 250, 0, 130, 256, 7, 90, 3, 255, 255, 255],
               "threshold": [256, 257, 255, 300, 0, 256, 2, 513, 1],
               "empty_runs": [0] * 37 + [5, 0, 0, 300, 0, 1] + [0] * 21}[pattern]
    n = 1203                                          # not a multiple of the block size
    ei = _degree_pattern_graph(n, degrees, seed=3, symmetric=pattern == "windows")
    o, c = oracle_and_cuda(f_in, f_out, aggrs, h, b, loops=loops, seed=5)
    torch.manual_seed(6)
    x, go = torch.randn(n, f_in), torch.randn(n, f_out)
    assert_close(run_both(o, c, x, ei, ei.to(DEV), go))
    rowptr, col, _ = to_adj_csr(ei, n)
    adj = egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), sparse_sizes=(n, n), is_sorted=True)
    assert_close(run_both(o, c, x, (rowptr, col, None), adj, go))


@pytest.mark.parametrize("loops,sigmoid,bias", list(itertools.product([True, False], [True, False], [True, False])))
def test_layer_flags(loops, sigmoid, bias):
    n = 500
    ei = random_graph(n, 2500, seed=41)
    o, c = oracle_and_cuda(32, 64, ["symnorm", "max", "std"], 4, 4, loops=loops, sigmoid=sigmoid, bias=bias, seed=2)
    torch.manual_seed(3)
    assert_close(run_both(o, c, torch.randn(n, 32), ei, ei.to(DEV), torch.randn(n, 64)))


def test_weighted_adjacency_without_symnorm():
    n = 700
    ei = random_graph(n, 4000, seed=51, hub=300)
    val = torch.rand(ei.size(1)) + 0.5
    rowptr, col, v = to_adj_csr(ei, n, val)
    o, c = oracle_and_cuda(24, 32, ["sum", "mean", "max", "min", "std"], 4, 2, seed=4)
    adj = egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), value=v.to(DEV), sparse_sizes=(n, n),
                                is_sorted=True)
    torch.manual_seed(5)
    assert_close(run_both(o, c, torch.randn(n, 24), (rowptr, col, v), adj, torch.randn(n, 32)))


def test_empty_and_tiny_graphs():
    o, c = oracle_and_cuda(8, 16, ["symnorm", "max", "std", "mean"], 4, 4, seed=1)
    x, go = torch.randn(6, 8), torch.randn(6, 16)
    for ei in (torch.zeros(2, 0, dtype=torch.long), torch.tensor([[1], [0]]), torch.tensor([[0, 0, 2], [0, 0, 2]])):
        assert_close(run_both(o, c, x, ei, ei.to(DEV), go))
    o2, c2 = oracle_and_cuda(8, 16, ["max", "std", "sum"], 4, 4, loops=False, seed=1)   # genuinely empty rows
    ei = torch.tensor([[1, 2], [0, 0]])
    assert_close(run_both(o2, c2, x, ei, ei.to(DEV), go))


def test_caching_reset_and_no_grad():
    n = 300
    ei = random_graph(n, 1500, seed=61).to(DEV)
    c = egc_b200.EGConv(16, 32, aggrs=["symnorm", "max"], num_heads=4, cached=True).to(DEV)
    x = torch.randn(n, 16, device=DEV)
    y1 = c(x, ei)
    assert c._cached_edge_index is not None
    other = random_graph(n, 900, seed=62).to(DEV)
    y2 = c(x, other)                                   # cached graph wins, like the reference (ref :129-141)
    assert torch.equal(y1, y2)
    c.reset_parameters()
    assert c._cached_edge_index is None and c._cached_adj_t is None
    with torch.no_grad():
        y3 = c(x, other)
    assert not y3.requires_grad and y3.shape == (n, 32)
    c.eval()
    assert torch.equal(c(x, other), y3)
    nc = egc_b200.EGConv(16, 32, aggrs=["symnorm"], num_heads=4, cached=False).to(DEV)
    nc(x, ei)
    assert nc._cached_edge_index is None


def test_three_layer_full_graph_step_replays_from_a_cuda_graph():
    """The reference's full-graph model body (mag/models.py:61-69: conv -> ReLU per layer, log_softmax + nll_loss)
    with cached structure: forward + backward captured once into ONE CUDA graph, replayed on new inputs."""
    n, classes = 5000, 40
    ei = random_graph(n, 60000, seed=12, hub=900)
    torch.manual_seed(2)
    convs = torch.nn.ModuleList([egc_b200.EGConv(64, 128, aggrs=["symnorm", "max", "std"], num_heads=4, num_bases=4, cached=True),
                                 egc_b200.EGConv(128, 128, aggrs=["symnorm", "max", "std"], num_heads=4, num_bases=4, cached=True),
                                 egc_b200.EGConv(128, classes, aggrs=["symnorm"], num_heads=8, num_bases=4, cached=True)]).to(DEV)
    params = list(convs.parameters())
    eid = ei.to(DEV)
    x_static = torch.randn(n, 64, device=DEV, requires_grad=True)
    y_static = torch.randint(0, classes, (n,), device=DEV)

    def step(x=None, y=None):
        x = x_static if x is None else x
        y = y_static if y is None else y
        h = x
        for i, c in enumerate(convs):
            h = c(h, eid)
            if i + 1 < len(convs):
                h = torch.relu(h)
        loss = torch.nn.functional.nll_loss(torch.log_softmax(h, dim=-1), y)
        return (loss,) + torch.autograd.grad(loss, [x] + params)

    before = egc_b200._lib.launch_count()
    step()
    per_step = egc_b200._lib.launch_count() - before
    graphed = egc_b200.GraphedStep(step, warmup=2)
    for it in range(3):
        gen = torch.Generator().manual_seed(70 + it)
        xf, yf = torch.randn(n, 64, generator=gen), torch.randint(0, classes, (n,), generator=gen)
        with torch.no_grad():
            x_static.copy_(xf.to(DEV))
            y_static.copy_(yf.to(DEV))
        counted = egc_b200._lib.launch_count()
        res = graphed.replay()
        torch.cuda.synchronize()
        assert egc_b200._lib.launch_count() == counted          # no eager launches: the step is one graph launch
        ref = step(xf.to(DEV).requires_grad_(True), yf.to(DEV))
        for a, b in zip(res, ref):
            assert rel_err(a, b.double().cpu()) < 1e-5
    assert per_step >= 3 * 8                                    # sanity: every layer contributes its kernels


@pytest.mark.parametrize("cfg", BLOCK_CONFIGS[:3], ids=lambda c: f"{c[0]}x{c[1]}-{'+'.join(c[2])}-h{c[3]}b{c[4]}")
def test_row_subset_launches_reproduce_the_whole_graph_call(cfg):
    """What the partitioned layer does: rows computed in two launches over complementary row subsets (the second one
    with the long-row plan) must give the bits of the single whole-graph launch - out, saved state and argmax positions."""
    from egc_b200 import functional as F
    f_in, f_out, aggrs, h, b = cfg
    n = 2500
    ei = _degree_pattern_graph(n, [3, 0, 40, 1, 300, 12, 7, 255, 256, 2, 90], seed=9, symmetric=False)
    g = egc_b200.GraphStructure.from_edge_index(ei.to(DEV), n, "symnorm" in aggrs, True)
    torch.manual_seed(4)
    bd = b * (f_out // h)
    bases = torch.randn(n, bd, device=DEV)
    weightings = torch.randn(n, h * len(aggrs) * b, device=DEV)
    bias = torch.randn(f_out, device=DEV)
    desc = F.make_desc(g, h, b, f_out // h, aggrs, False)
    whole = F.aggregate_combine(desc, g, bases, weightings, bias, want_saved=True)
    pick = torch.rand(n, generator=torch.Generator().manual_seed(1)) < 0.6
    first = torch.nonzero(pick).flatten().to(torch.int32).to(DEV)
    second = torch.nonzero(~pick).flatten().to(torch.int32).to(DEV)
    outs = F.alloc_aggregate_outputs(desc, DEV, want_out=True, want_saved=True)
    for t in outs:
        if t is not None:
            t.fill_(0)
    F.aggregate_combine(desc, g, bases, weightings, bias, row_subset=first, use_plan=False, outputs=outs)
    F.aggregate_combine(desc, g, bases, weightings, bias, row_subset=second, use_plan=True, outputs=outs)
    for a, ref in zip(outs, whole):
        assert (a is None) == (ref is None)
        if a is not None:
            assert torch.equal(a, ref)
