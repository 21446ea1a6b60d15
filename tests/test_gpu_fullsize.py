"""Parity at the sizes BASELINE.json names (the oracle comparisons elsewhere run on graphs of <= 40 k nodes):

  * cfg2: the ogbn-arxiv-shaped EGC-M layer, forward + EVERY gradient against the fp64 CPU oracle on the full
    `bench.synth_graph("arxiv")`, both input kinds;
  * cfg4: the ogbn-mag-shaped EGC-S layer, forward rows + gradients of a loss over a 4 % sample of the target rows
    (the oracle aggregates only those rows - with the global symnorm weights - so it finishes in seconds);
  * the tcgen05 projection kernels (`k_project_tc`, `k_wgrad_tc`) called directly through the C ABI with
    GEMM_3XTF32 / GEMM_TF32 against fp64, with the bar each one is held to written next to the measured error.
"""
import pytest
import torch

import egc_b200
from bench import WORKLOADS, synth_graph, to_adj_t
from egc_b200 import _lib
from egc_b200.functional import project
from oracle import restatement as R
from tests.util import rel_err, restrict_rows

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5                       # BASELINE.json north_star: features and gradients, max|a-b| / max|b| in fp32
TOL_3XTF32 = 1e-5                # stated tolerance of the default tensor-core path (3-term TF32 split)
TOL_TF32 = 3e-3                  # stated tolerance of the opt-in single-pass TF32 path (10-bit mantissa operands)


def _layer_pair(w, seed=0):
    torch.manual_seed(seed)
    o = R.EGConvOracle(w["f_in"], w["f_out"], aggrs=w["aggrs"], num_heads=w["heads"], num_bases=w["bases"])
    with torch.no_grad():
        o.bias.uniform_(-0.5, 0.5)
    c = egc_b200.EGConv(w["f_in"], w["f_out"], aggrs=w["aggrs"], num_heads=w["heads"], num_bases=w["bases"])
    c.load_state_dict(o.state_dict())
    return o, c.to(DEV)


def _oracle_grads(o, x, graph_in, go, dtype):
    od = R.EGConvOracle(o.in_channels, o.out_channels, aggrs=o.aggregators, num_heads=o.num_heads,
                        num_bases=o.num_bases).to(dtype)
    od.load_state_dict({k: v.to(dtype) for k, v in o.state_dict().items()})
    xo = x.to(dtype).requires_grad_(True)
    out = od(xo, graph_in)
    names = [n for n, _ in od.named_parameters()]
    grads = torch.autograd.grad(out, [xo] + list(od.parameters()), go.to(dtype))
    return out.detach(), dict(zip(["x"] + names, grads))


@pytest.mark.parametrize("kind", ["edge_index", "adj_t"])
def test_arxiv_full_size_layer_vs_oracle(kind):
    w = WORKLOADS["arxiv"]
    n, ei = synth_graph("arxiv", 0)
    o, c = _layer_pair(w)
    gen = torch.Generator().manual_seed(3)
    x, go = torch.randn(n, w["f_in"], generator=gen), torch.randn(n, w["f_out"], generator=gen)
    if kind == "edge_index":
        g_cpu, g_gpu = ei, ei.to(DEV)
    else:
        rowptr, col = to_adj_t(ei, n)
        g_cpu = (rowptr, col, None)
        g_gpu = egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), sparse_sizes=(n, n), is_sorted=True)
    out64, g64 = _oracle_grads(o, x, g_cpu, go, torch.float64)
    out32, g32 = _oracle_grads(o, x, g_cpu, go, torch.float32)
    xc = x.to(DEV).requires_grad_(True)
    out = c(xc, g_gpu)
    names = [k for k, _ in c.named_parameters()]
    grads = dict(zip(["x"] + names, torch.autograd.grad(out, [xc] + list(c.parameters()), go.to(DEV))))
    e = rel_err(out, out64)
    assert e < TOL, f"out: {e:.3e}"
    report = {"out": e}
    for k in grads:
        bar = max(TOL, 4.0 * rel_err(g32[k], g64[k]))           # std's fp32 cancellation noise, see test_gpu_parity.tolerance
        e = rel_err(grads[k], g64[k])
        report[k] = (e, bar)
        assert e < bar, f"grad {k}: {e:.3e} >= bar {bar:.3e}"
    print("full-size arxiv", kind, report)


def test_mag_full_size_sampled_rows_vs_oracle():
    w = WORKLOADS["mag"]
    n, ei = synth_graph("mag", 0)
    rowptr, col = to_adj_t(ei, n)
    o, c = _layer_pair(w)
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(n, w["f_in"], generator=gen)
    keep = torch.rand(n, generator=gen) < 0.04
    keep[:64] = True                                            # the Zipf hubs' neighbours and a dense head of rows
    go = torch.randn(n, w["f_out"], generator=gen) * keep.view(-1, 1)
    g_full = R.graph_from_csr(rowptr, col, None, n, True, True, False)
    g_s = restrict_rows(g_full, keep)

    def oracle(dtype):
        p = {k: v.to(dtype).requires_grad_(True) for k, v in o.state_dict().items()}
        xo = x.to(dtype).requires_grad_(True)
        gs = g_s if dtype == torch.float32 else restrict_rows(
            R.graph_from_csr(rowptr, col, None, n, True, True, False, dtype), keep)
        out = R.egconv_forward(xo, gs, p["bases_weight"], p["comb_weight.weight"], p["comb_weight.bias"], p["bias"],
                               w["aggrs"], w["heads"])
        grads = torch.autograd.grad(out, [xo] + list(p.values()), go.to(dtype))
        return out.detach(), dict(zip(["x"] + list(p.keys()), grads))

    out64, g64 = oracle(torch.float64)
    out32, g32 = oracle(torch.float32)
    xc = x.to(DEV).requires_grad_(True)
    adj = egc_b200.SparseTensor(rowptr=rowptr.to(DEV), col=col.to(DEV), sparse_sizes=(n, n), is_sorted=True)
    out = c(xc, adj)
    names = [k for k, _ in c.named_parameters()]
    grads = dict(zip(["x"] + names, torch.autograd.grad(out, [xc] + list(c.parameters()), go.to(DEV))))
    # structure first: bit-exact CSR + symnorm weights at full size
    g = c._prepare(xc, adj) if c.cached else egc_b200.GraphStructure.from_csr(rowptr.to(DEV), col.to(DEV), None, n, True, True)
    assert torch.equal(g.rowptr.cpu().long(), g_full.rowptr) and torch.equal(g.col.cpu().long(), g_full.col)
    assert torch.equal(g.val_sym.cpu(), g_full.val_sym)
    e = rel_err(out[keep.to(DEV)], out64[keep])
    assert e < TOL, f"out rows: {e:.3e}"
    report = {"out": e}
    for k in grads:
        bar = max(TOL, 4.0 * rel_err(g32[k], g64[k]))
        e = rel_err(grads[k], g64[k])
        report[k] = (e, bar)
        assert e < bar, f"grad {k}: {e:.3e} >= bar {bar:.3e}"
    print("full-size mag (4 % of the target rows)", report)


@pytest.mark.parametrize("algo,bar", [(_lib.GEMM_3XTF32, TOL_3XTF32), (_lib.GEMM_TF32, TOL_TF32)], ids=["3xtf32", "tf32"])
@pytest.mark.parametrize("n,f_in,bd,hab", [(1000, 128, 128, 48), (169_343, 128, 128, 48), (4097, 128, 64, 32),
                                           (777, 100, 52, 32), (130, 24, 176, 96), (333, 352, 64, 32)])
@pytest.mark.parametrize("sigmoid", [False, True])
def test_tensor_core_projection_vs_fp64(algo, bar, n, f_in, bd, hab, sigmoid):
    """`k_project_tc` (forward and d_x) and `k_wgrad_tc` through egc_project_fwd / egc_project_bwd with the algorithm
    forced (no AUTO fallback): the shapes of test_project_fwd_bwd plus the BASELINE ones."""
    lib = egc_b200.load()
    torch.manual_seed(0)
    x, wb = torch.randn(n, f_in), torch.randn(f_in, bd) * 0.1
    wc, bc = torch.randn(hab, f_in) * 0.1, torch.randn(hab)
    before = _lib.launch_count()
    try:
        bases, wts = project(x.to(DEV), wb.to(DEV), wc.to(DEV), bc.to(DEV), sigmoid, algo)
    except egc_b200.EGCError as err:                            # a forced tensor-core algorithm never falls back silently
        assert "does not support" in str(err)
        pytest.skip(f"shape outside the tensor-core path: {err}")
    b_o, w_o = R.project(x.double(), wb.double(), wc.double(), bc.double(), sigmoid)
    errs = {"bases": rel_err(bases, b_o), "weightings": rel_err(wts, w_o)}
    d_bases, d_lin = torch.randn(n, bd), torch.randn(n, hab)
    outs = [torch.empty(s, device=DEV) for s in ((n, f_in), (f_in, bd), (hab, f_in), (hab,))]
    nbytes = lib.egc_project_bwd_workspace_bytes(n, f_in, bd, hab)
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=DEV)
    P = _lib.ptr
    dev_in = [t.to(DEV) for t in (x, wb, wc, d_bases, d_lin)]
    _lib.check(lib.egc_project_bwd(*[P(t) for t in dev_in], n, f_in, bd, hab, *[P(t) for t in outs], algo, P(ws), nbytes,
                                   torch.cuda.current_stream().cuda_stream), "egc_project_bwd")
    torch.cuda.synchronize()
    assert _lib.launch_count() > before
    xd, db, dl = x.double(), d_bases.double(), d_lin.double()
    errs["d_x"] = rel_err(outs[0], db @ wb.double().t() + dl @ wc.double())
    errs["d_bases_weight"] = rel_err(outs[1], xd.t() @ db)
    errs["d_comb_weight"] = rel_err(outs[2], dl.t() @ xd)
    errs["d_comb_bias"] = rel_err(outs[3], dl.sum(0))
    print(f"tensor-core projection n={n} f_in={f_in} bd={bd} hab={hab} sigmoid={sigmoid} bar={bar:g}: {errs}")
    for k, e in errs.items():
        assert e < bar, f"{k}: {e:.3e} >= {bar:g}"


@pytest.mark.parametrize("n,f_in,bd,hab", [(2001, 128, 64, 224), (5000, 128, 32, 288), (999, 64, 100, 544)])
def test_wide_parameter_gradient_runs_on_the_tensor_cores(n, f_in, bd, hab):
    """More than 256 accumulator columns (REGConv's paper type: root + every relation's combination weights in one
    projection, rmag/models.py:100-146): `k_wgrad_mn` is launched once per <= 256-column range instead of falling back to
    the fp32 FFMA kernel; results vs fp64 at the 3xTF32 bar."""
    lib = egc_b200.load()
    torch.manual_seed(1)
    x, d_bases, d_lin = torch.randn(n, f_in), torch.randn(n, bd), torch.randn(n, hab)
    wb, wc = torch.randn(f_in, bd) * 0.1, torch.randn(hab, f_in) * 0.1
    outs = [torch.empty(s, device=DEV) for s in ((f_in, bd), (hab, f_in), (hab,))]
    nbytes = lib.egc_project_bwd_workspace_bytes(n, f_in, bd, hab)
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=DEV)
    P = _lib.ptr
    dev_in = [t.to(DEV) for t in (x, wb, wc, d_bases, d_lin)]
    _lib.profile_enable(True)
    try:
        rc = lib.egc_project_bwd(*[P(t) for t in dev_in], n, f_in, bd, hab, None, *[P(t) for t in outs], _lib.GEMM_3XTF32,
                                 P(ws), nbytes, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        prof = _lib.profile_collect()
    finally:
        _lib.profile_enable(False)
    if rc != 0:
        assert b"does not support" in lib.egc_last_error_string()
        pytest.skip("shape outside the tensor-core projection")
    assert prof.get("k_wgrad_tc", (0, 0))[0] == 1 + -(-(hab - (256 - 32 * -(-bd // 32)) // 32 * 32) // 256), prof
    xd = x.double()
    errs = {"d_bases_weight": rel_err(outs[0], xd.t() @ d_bases.double()), "d_comb_weight": rel_err(outs[1], d_lin.double().t() @ xd),
            "d_comb_bias": rel_err(outs[2], d_lin.double().sum(0))}
    print(f"wide wgrad n={n} f_in={f_in} bd={bd} hab={hab}: {errs} {prof}")
    for k, e in errs.items():
        assert e < TOL_3XTF32, f"{k}: {e:.3e}"
