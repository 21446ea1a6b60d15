"""The full-graph layer stack `EGC` (ref experiments/mag/models.py:16-69): the CPU restatement against golden vectors
from the UNMODIFIED reference class (oracle/make_golden_stack.py), and the sm_100a stack against both."""
import pytest
import torch
import torch.nn.functional as F

from oracle import restatement as R
from tests.util import golden_cases, load_golden, rel_err

STACKS = golden_cases("egc_stack_")


def _oracle(rec, dtype, dropout=0.0):
    m = R.EGCOracle(rec["hidden"], rec["layers"], dropout, rec["heads"], rec["bases"], rec["aggrs"]).to(dtype)
    m.load_state_dict({k: v.to(dtype) for k, v in rec["state_dict"].items()})
    return m


@pytest.mark.parametrize("name", STACKS)
def test_stack_restatement_matches_reference_golden(name):
    rec = load_golden(name)
    for tag, dt, tol in (("f64", torch.float64, 1e-11), ("f32", torch.float32, 2e-5)):
        m = _oracle(rec, dt).train()
        x = rec["x"].to(dt).requires_grad_(True)
        out = m(x, (rec["adj_rowptr"], rec["adj_col"], None))
        loss = F.nll_loss(out[rec["train_idx"]], rec["y"][rec["train_idx"]])
        params = list(m.named_parameters())
        grads = torch.autograd.grad(loss, [x] + [p for _, p in params])
        assert rel_err(out, rec[f"out_{tag}"]) < tol
        assert abs(float(loss.detach()) - float(rec[f"loss_{tag}"])) < tol * max(1.0, abs(float(rec[f"loss_{tag}"])))
        assert rel_err(grads[0], rec[f"grad_x_{tag}"]) < tol * 10
        for (pn, _), g in zip(params, grads[1:]):
            assert rel_err(g, rec[f"grad_{pn}_{tag}"]) < tol * 10, pn
    # eval mode switches dropout off (ref :65 `training=self.training`)
    m = _oracle(rec, torch.float32, dropout=0.5).eval()
    with torch.no_grad():
        assert rel_err(m(rec["x"], (rec["adj_rowptr"], rec["adj_col"], None)), rec["out_eval_dropout_f32"]) < 2e-5


def test_stack_state_dict_keys_and_param_count():
    """Shape known-answer: state_dict layout `convs.{i}.{bases_weight, comb_weight.weight, comb_weight.bias, bias}` and
    the parameter count that follows from ref optimized_layers.py:105-115 for dims 128 -> 64 -> 64 -> 352."""
    import egc_b200
    m = egc_b200.EGC(64, 3, 0.5, 8, 4, ["symnorm"])
    o = R.EGCOracle(64, 3, 0.5, 8, 4, ["symnorm"])
    assert list(m.state_dict().keys()) == list(o.state_dict().keys())
    assert [tuple(v.shape) for v in m.state_dict().values()] == [tuple(v.shape) for v in o.state_dict().values()]
    dims = [128, 64, 64, 352]
    want = sum(f * (g // 8) * 4 + (8 * 4) * f + 8 * 4 + g for f, g in zip(dims[:-1], dims[1:]))
    assert sum(p.numel() for p in m.parameters()) == want
    with pytest.raises(ValueError):
        egc_b200.EGC(64, 1, 0.0, 8, 4, ["symnorm"])


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", STACKS)
@pytest.mark.parametrize("kind", ["adj_t", "edge_index"])
@pytest.mark.parametrize("gemm", ["fp32_simt", "auto"])
def test_stack_cuda_matches_reference_golden(name, kind, gemm):
    """egc_b200.EGC (CUDA kernels) against the unmodified reference stack: log-probabilities, loss, all gradients.
    Bars (the one used is part of the assertion message):
      * exact-fp32 projections (GEMM_FP32_SIMT): 1e-5 on the output; on the gradients 1e-5 or 4x the reference's own
        fp32-vs-fp64 error where that is larger (std's cancellation) - the single-layer bar holds through 3 layers;
      * default tensor-core projections (3xTF32, ~2e-6 per projection): 1e-5 x layers."""
    import egc_b200
    from egc_b200 import _lib
    rec = load_golden(name)
    dev = "cuda:0"
    algo = _lib.GEMM_FP32_SIMT if gemm == "fp32_simt" else _lib.GEMM_AUTO
    base = 1e-5 if gemm == "fp32_simt" else 1e-5 * rec["layers"]
    m = egc_b200.EGC(rec["hidden"], rec["layers"], 0.0, rec["heads"], rec["bases"], rec["aggrs"], gemm_algo=algo).to(dev).train()
    m.load_state_dict(rec["state_dict"])
    x = rec["x"].to(dev).requires_grad_(True)
    if kind == "adj_t":
        gin = egc_b200.SparseTensor(rowptr=rec["adj_rowptr"].to(dev), col=rec["adj_col"].to(dev),
                                    sparse_sizes=(rec["n"], rec["n"]), is_sorted=True)
    else:
        gin = rec["edge_index"].to(dev)
    out = m(x, gin)
    idx = rec["train_idx"].to(dev)
    loss = F.nll_loss(out[idx], rec["y"].to(dev)[idx])
    params = list(m.named_parameters())
    grads = torch.autograd.grad(loss, [x] + [p for _, p in params])
    e = rel_err(out, rec["out_f64"])
    assert e < base, f"out: {e:.3e} >= bar {base:.1e}"
    assert abs(float(loss.detach()) - float(rec["loss_f64"])) < base * abs(float(rec["loss_f64"]))
    for (pn, g) in zip(["x"] + [n for n, _ in params], grads):
        ref64, ref32 = rec[f"grad_{pn}_f64"], rec[f"grad_{pn}_f32"]
        bar = max(base, 4.0 * rel_err(ref32, ref64))
        e = rel_err(g, ref64)
        assert e < bar, f"grad {pn}: {e:.3e} >= bar {bar:.3e}"
    m2 = egc_b200.EGC(rec["hidden"], rec["layers"], 0.5, rec["heads"], rec["bases"], rec["aggrs"], gemm_algo=algo).to(dev).eval()
    m2.load_state_dict(rec["state_dict"])
    with torch.no_grad():
        assert rel_err(m2(rec["x"].to(dev), gin), rec["out_eval_dropout_f32"]) < base


@pytest.mark.gpu
def test_stack_dropout_is_active_in_training_mode():
    import egc_b200
    rec = load_golden(STACKS[0])
    dev = "cuda:0"
    m = egc_b200.EGC(rec["hidden"], rec["layers"], 0.5, rec["heads"], rec["bases"], rec["aggrs"]).to(dev).train()
    m.load_state_dict(rec["state_dict"])
    gin = rec["edge_index"].to(dev)
    with torch.no_grad():
        a, b = m(rec["x"].to(dev), gin), m(rec["x"].to(dev), gin)
    assert rel_err(a, b) > 1e-3            # two dropout draws differ
