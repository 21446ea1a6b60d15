"""Generate tests/golden/arxivnet_*.pt from the UNMODIFIED reference model `experiments/arxiv/norm_models.py::EgcArxivNet`
(Linear embed -> [EfficientGraphConv -> BatchNorm1d -> ReLU -> dropout -> + identity] x L -> Linear -> log_softmax),
imported through oracle/shims (run in the build container):

    python -m oracle.make_golden_arxivnet

`norm_models.py` also imports the baseline layers of PyG (GATConv, GCNConv, ...) at module level; they are not on the
path under test and not in the shims, so placeholder names are injected into the shim module before the import - the
reference file itself is untouched.  Stored per case: inputs, the full state_dict (BatchNorm running statistics
perturbed so that eval mode is not the identity), eval-mode output / loss / gradients (frozen statistics) and
training-mode ones with dropout 0 (batch statistics), in fp32 and fp64."""
import importlib
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402
from oracle.make_golden import make_graph  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")
# name, N, E, hidden, layers, heads, bases, aggrs, residual         (ref arxiv/configs.py:325-345: egc_s / egc_m settings)
CASES = [
    ("arxivnet_s", 140, 900, 64, 3, 8, 4, ["symadd"], True),
    ("arxivnet_m", 120, 800, 64, 2, 4, 4, ["symadd", "max", "mean"], True),
    ("arxivnet_s_plain", 100, 600, 32, 2, 4, 4, ["symadd"], False),
]


def load_models():
    rl.load()
    tg_nn = importlib.import_module("torch_geometric.nn")
    for name in ("GATConv", "GATv2Conv", "GCNConv", "GINConv", "PNAConv", "SAGEConv"):
        if not hasattr(tg_nn, name):
            setattr(tg_nn, name, type(name, (), {}))          # never instantiated by EgcArxivNet
    return importlib.import_module("experiments.arxiv.norm_models")


def main():
    models = load_models()
    for name, n, e, hidden, layers, heads, bases, aggrs, residual in CASES:
        gen = torch.Generator().manual_seed(sum(ord(c) for c in name))
        ei = make_graph(n, e, 0, gen)
        x = torch.randn(n, models.NUM_FEATURES, generator=gen)
        y = torch.randint(0, models.NUM_CLASSES, (n,), generator=gen)
        train_idx = torch.randperm(n, generator=gen)[: n // 2]
        torch.manual_seed(77)
        model = models.EgcArxivNet(hidden, layers, 0.0, residual, heads=heads, bases=bases, softmax=False, aggrs=aggrs)
        with torch.no_grad():
            for bn in model.bns:
                bn.running_mean.normal_(0, 0.3, generator=gen)
                bn.running_var.uniform_(0.5, 2.0, generator=gen)
                bn.weight.uniform_(0.5, 1.5, generator=gen)
                bn.bias.normal_(0, 0.2, generator=gen)
            for conv in model.convs:
                conv.bias.uniform_(-0.3, 0.3, generator=gen)
        rec = {"name": name, "n": n, "hidden": hidden, "layers": layers, "heads": heads, "bases": bases, "aggrs": aggrs,
               "residual": residual, "edge_index": ei, "x": x, "y": y, "train_idx": train_idx,
               "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
               "num_params": sum(p.numel() for p in model.parameters())}
        for mode in ("eval", "train"):
            for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
                m = models.EgcArxivNet(hidden, layers, 0.0, residual, heads=heads, bases=bases, softmax=False, aggrs=aggrs).to(dt)
                m.load_state_dict({k: (v.to(dt) if v.is_floating_point() else v) for k, v in rec["state_dict"].items()})
                m.train(mode == "train")
                xx = x.to(dt).requires_grad_(True)
                out = m(xx, ei)
                loss = F.nll_loss(out[train_idx], y[train_idx])
                params = list(m.named_parameters())
                grads = torch.autograd.grad(loss, [xx] + [p for _, p in params])
                rec[f"{mode}_out_{tag}"], rec[f"{mode}_loss_{tag}"], rec[f"{mode}_grad_x_{tag}"] = out.detach(), loss.detach(), grads[0]
                for (pn, _), g in zip(params, grads[1:]):
                    rec[f"{mode}_grad_{pn}_{tag}"] = g
        path = os.path.join(OUT_DIR, f"{name}.pt")
        torch.save(rec, path)
        print(f"{name:18s} params {rec['num_params']}  eval loss {float(rec['eval_loss_f64']):.6f}  train loss "
              f"{float(rec['train_loss_f64']):.6f}  {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
