"""Fused epilogue of the aggregation kernel: eval-mode BatchNorm folded to an affine map, ReLU and
the post-activation residual of the reference's normalised stacks (/root/reference/experiments/arxiv/norm_models.py:33-40,
zinc/models.py:60-74), and REGConv's accumulate-into-output (rmag/models.py:146)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import restatement as R
from tests.util import random_graph, rel_err

DEV = "cuda:0"


def test_fold_batchnorm_equals_eval_mode_batchnorm():
    from egc_b200.stack import fold_batchnorm
    torch.manual_seed(0)
    for affine in (True, False):
        bn = torch.nn.BatchNorm1d(24, affine=affine).double()
        with torch.no_grad():
            bn.running_mean.normal_()
            bn.running_var.uniform_(0.2, 3.0)
            if affine:
                bn.weight.normal_()
                bn.bias.normal_()
        bn.eval()
        y = torch.randn(50, 24, dtype=torch.float64)
        scale, shift = fold_batchnorm(bn)
        assert scale.dtype == torch.float32 and not scale.requires_grad
        assert rel_err(y * scale.double() + shift.double(), bn(y)) < 1e-6
    with pytest.raises(ValueError):
        fold_batchnorm(torch.nn.BatchNorm1d(8, track_running_stats=False))


def test_block_rejects_a_residual_with_changing_width():
    import egc_b200
    with pytest.raises(ValueError):
        egc_b200.EGCBlock(egc_b200.EGConv(16, 32, num_heads=4), residual=True)


EPI_CONFIGS = [  # f_in, f_out, aggrs, heads, bases
    (128, 128, ["symnorm", "max", "std"], 4, 4),              # cfg2 shape: row-block kernel, 128-bit path
    (64, 64, ["sum", "mean", "min", "var"], 4, 4),            # G = 16
    (32, 40, ["max", "std", "sum"], 4, 3),                    # B*D = 30: general kernel, scalar epilogue
    (128, 352, ["mean"], 8, 4),                               # two feature passes
]


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", EPI_CONFIGS, ids=lambda c: f"{c[0]}x{c[1]}-{'+'.join(c[2])}")
@pytest.mark.parametrize("parts", ["affine", "affine+relu", "residual", "affine+relu+residual", "relu+residual"])
def test_fused_epilogue_equals_the_separate_ops(cfg, parts):
    """conv(x, g, relu=, scale=, shift=, residual=) against relu(conv(x, g) * scale + shift) + residual built from torch
    ops on the same layer: the forward differs by one fused-multiply-add rounding, the gradients of x, every parameter and
    the residual agree to 1e-5 (deterministic routing so min/max do not add atomic-order noise)."""
    import egc_b200
    f_in, f_out, aggrs, h, b = cfg
    n = 3000
    ei = random_graph(n, 30000, seed=91, hub=600).to(DEV)
    torch.manual_seed(13)
    c = egc_b200.EGConv(f_in, f_out, aggrs=aggrs, num_heads=h, num_bases=b).to(DEV)
    with torch.no_grad():
        c.bias.uniform_(-0.5, 0.5)
    c.deterministic = True
    x, go = torch.randn(n, f_in, device=DEV), torch.randn(n, f_out, device=DEV)
    scale = torch.empty(f_out, device=DEV).uniform_(0.3, 2.0) * torch.where(torch.rand(f_out, device=DEV) < 0.2, -1.0, 1.0)
    shift = torch.randn(f_out, device=DEV) * 0.3
    res = torch.randn(n, f_out, device=DEV)
    use_affine, use_relu, use_res = "affine" in parts, "relu" in parts, "residual" in parts

    xa, ra = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    ya = c(xa, ei)
    if use_affine:
        ya = ya * scale + shift
    if use_relu:
        ya = torch.relu(ya)
    if use_res:
        ya = ya + ra
    ins_a = [xa] + list(c.parameters()) + ([ra] if use_res else [])
    ga = torch.autograd.grad(ya, ins_a, go)

    xb, rb = x.clone().requires_grad_(True), res.clone().requires_grad_(True)
    yb = c(xb, ei, relu=use_relu, scale=scale if use_affine else None, shift=shift if use_affine else None,
           residual=rb if use_res else None)
    ins_b = [xb] + list(c.parameters()) + ([rb] if use_res else [])
    gb = torch.autograd.grad(yb, ins_b, go)

    assert rel_err(yb, ya) < 1e-6
    if not use_affine:
        assert torch.equal(ya, yb)                            # no fma involved: the same bits
    for name, u, v in zip(["x"] + [k for k, _ in c.named_parameters()] + ["residual"], ga, gb):
        assert rel_err(v, u) < 1e-5, f"grad {name}: {rel_err(v, u):.3e}"
    with torch.no_grad():                                     # inference path (nothing saved)
        yc = c(x, ei, relu=use_relu, scale=scale if use_affine else None, shift=shift if use_affine else None,
               residual=res if use_res else None)
    assert torch.equal(yc, yb.detach())


@pytest.mark.gpu
def test_fused_epilogue_validates_its_inputs():
    import egc_b200
    n = 200
    ei = random_graph(n, 1000, seed=92).to(DEV)
    c = egc_b200.EGConv(32, 32, aggrs=["sum"], num_heads=4).to(DEV)
    x = torch.randn(n, 32, device=DEV)
    s = torch.ones(32, device=DEV)
    with pytest.raises(ValueError):
        c(x, ei, scale=s)                                     # scale without shift
    with pytest.raises(ValueError):
        c(x, ei, scale=s[:16], shift=s[:16])                  # wrong width
    with pytest.raises(ValueError):
        c(x, ei, residual=x[:10])                             # wrong row count
    with pytest.raises(ValueError):
        c(x, ei, scale=s.clone().requires_grad_(True), shift=s)   # constants only


@pytest.mark.gpu
@pytest.mark.parametrize("residual", [False, True])
def test_block_stack_in_eval_mode_matches_the_fp64_oracle(residual):
    """Two `EGCBlock`s (conv -> BN -> ReLU -> dropout -> + identity, arxiv/norm_models.py:33-40) in eval mode - the whole
    tail inside the aggregation kernel - against the oracle layer in fp64 followed by torch's own eval-mode BatchNorm,
    ReLU and add: output and the gradients of the input and of every conv parameter."""
    import egc_b200
    n, f = 4000, 64
    ei = random_graph(n, 36000, seed=93, hub=500)
    aggrs = ["symnorm", "max", "std"]
    torch.manual_seed(14)
    blocks = torch.nn.ModuleList([egc_b200.EGCBlock(egc_b200.EGConv(f, f, aggrs=aggrs, num_heads=4, num_bases=4), 0.3, residual)
                                  for _ in range(2)])
    oracles = []
    for blk in blocks:
        with torch.no_grad():
            blk.bn.running_mean.normal_(0, 0.3)
            blk.bn.running_var.uniform_(0.5, 2.0)
            blk.bn.weight.uniform_(0.5, 1.5)
            blk.bn.bias.normal_(0, 0.2)
        o = R.EGConvOracle(f, f, aggrs=aggrs, num_heads=4, num_bases=4).double()
        o.load_state_dict({k: v.double() for k, v in blk.conv.state_dict().items()})
        bn = torch.nn.BatchNorm1d(f).double()
        bn.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in blk.bn.state_dict().items()})
        oracles.append((o, bn.eval()))
    blocks = blocks.to(DEV).eval()
    x, go = torch.randn(n, f), torch.randn(n, f)

    xo = x.double().requires_grad_(True)
    y = xo
    for o, bn in oracles:
        z = F.relu(bn(o(y, ei)))                              # dropout is the identity in eval mode
        y = z + y if residual else z
    params_o = [p for o, _ in oracles for p in o.parameters()]
    g_o = torch.autograd.grad(y, [xo] + params_o, go.double())

    xc = x.to(DEV).requires_grad_(True)
    yc = xc
    gin = ei.to(DEV)
    for blk in blocks:
        yc = blk(yc, gin)
    params_c = [p for blk in blocks for p in blk.conv.parameters()]
    g_c = torch.autograd.grad(yc, [xc] + params_c, go.to(DEV))
    assert rel_err(yc, y) < 2e-5
    for i, (a, b) in enumerate(zip(g_c, g_o)):
        assert rel_err(a, b) < 5e-5, f"gradient {i}: {rel_err(a, b):.3e}"     # std's fp32 cancellation through two layers


@pytest.mark.gpu
def test_block_in_training_mode_uses_batch_statistics():
    import egc_b200
    n, f = 1000, 32
    ei = random_graph(n, 6000, seed=94).to(DEV)
    torch.manual_seed(15)
    blk = egc_b200.EGCBlock(egc_b200.EGConv(f, f, aggrs=["sum", "max"], num_heads=4), 0.0, True).to(DEV).train()
    x = torch.randn(n, f, device=DEV)
    before = blk.bn.running_mean.clone()
    y = blk(x, ei)
    want = F.relu(F.batch_norm(blk.conv(x, ei), None, None, blk.bn.weight, blk.bn.bias, True)) + x
    assert rel_err(y, want) < 1e-6 and not torch.equal(before, blk.bn.running_mean)
