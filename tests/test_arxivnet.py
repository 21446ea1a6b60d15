"""`egc_b200.EgcArxivNet` against golden vectors produced by the UNMODIFIED reference model
(/root/reference/experiments/arxiv/norm_models.py::EgcArxivNet, oracle/make_golden_arxivnet.py): the stack around the
layer - embed, BatchNorm, ReLU, residual, classifier, log_softmax.  In eval mode the EGC-S cases
run every block as ONE aggregation kernel (folded BatchNorm + ReLU + residual in its epilogue)."""
import os

import pytest
import torch
import torch.nn.functional as F

from tests.util import rel_err

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["arxivnet_s", "arxivnet_m", "arxivnet_s_plain"]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def build(rec):
    import egc_b200
    m = egc_b200.EgcArxivNet(rec["hidden"], rec["layers"], 0.0, rec["residual"], heads=rec["heads"], bases=rec["bases"],
                             softmax=False, aggrs=rec["aggrs"])
    m.load_state_dict(rec["state_dict"])           # strict: every key of the reference's state_dict, nothing else
    return m


@pytest.mark.parametrize("name", CASES)
def test_state_dict_layout_and_parameter_count_match_the_reference(name):
    rec = load_golden(name)
    m = build(rec)
    assert list(m.state_dict().keys()) == list(rec["state_dict"].keys())
    assert sum(p.numel() for p in m.parameters()) == rec["num_params"]
    assert m._fusable() is False                   # training mode after construction
    assert m.eval()._fusable() == (rec["aggrs"] == ["symadd"])


@pytest.mark.parametrize("mode", ["eval", "train"])
@pytest.mark.parametrize("name", CASES)
def test_oracle_restatement_matches_reference_golden(name, mode):
    """The CPU restatement of the model (oracle/restatement.py::arxivnet_forward) against the unmodified reference's
    output, loss and input gradient in fp64: pins what the goldens mean independently of the CUDA path."""
    from oracle import restatement as R
    rec = load_golden(name)
    state = {k: (v.double() if v.is_floating_point() else v) for k, v in rec["state_dict"].items()}
    x = rec["x"].double().requires_grad_(True)
    out = R.arxivnet_forward(x, rec["edge_index"], state, rec["layers"], rec["heads"], rec["bases"], rec["aggrs"],
                             rec["residual"], training=(mode == "train"))
    loss = F.nll_loss(out[rec["train_idx"]], rec["y"][rec["train_idx"]])
    (gx,) = torch.autograd.grad(loss, [x])
    assert rel_err(out, rec[f"{mode}_out_f64"]) < 1e-11
    assert abs(float(loss) - float(rec[f"{mode}_loss_f64"])) < 1e-11
    assert rel_err(gx, rec[f"{mode}_grad_x_f64"]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["eval", "train"])
@pytest.mark.parametrize("name", CASES)
def test_arxivnet_cuda_matches_reference_golden(name, mode):
    rec = load_golden(name)
    dev = "cuda:0"
    m = build(rec).to(dev)
    m.train(mode == "train")
    fused = m._fusable()
    x = rec["x"].to(dev).requires_grad_(True)
    out = m(x, rec["edge_index"].to(dev))
    loss = F.nll_loss(out[rec["train_idx"].to(dev)], rec["y"].to(dev)[rec["train_idx"].to(dev)])
    named = list(m.named_parameters())
    grads = torch.autograd.grad(loss, [x] + [p for _, p in named], allow_unused=True)
    layers = rec["layers"]

    def bar(key):                                   # 1e-5 per layer, or 4x the reference's own fp32-vs-fp64 error
        return max(1e-5 * layers, 4.0 * rel_err(rec[f"{mode}_{key}_f32"], rec[f"{mode}_{key}_f64"]))

    report = {"out": (rel_err(out, rec[f"{mode}_out_f64"]), bar("out"))}
    assert report["out"][0] < report["out"][1], report
    assert abs(float(loss) - float(rec[f"{mode}_loss_f64"])) < 1e-5 * max(1.0, abs(float(rec[f"{mode}_loss_f64"])))
    for (pn, _), g in zip([("x", None)] + named, grads):
        if g is None:                              # fused eval path: the folded BatchNorm affine is a constant
            assert fused and pn.startswith("bns."), pn
            continue
        e, b = rel_err(g, rec[f"{mode}_grad_{pn}_f64"]), bar(f"grad_{pn}")
        report[pn] = (e, b)
        assert e < b, f"{name} {mode} grad {pn}: {e:.3e} >= {b:.3e}"
    print(name, mode, "fused" if fused else "unfused", {k: f"{v[0]:.1e}" for k, v in report.items()})
