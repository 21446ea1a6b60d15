"""Generate tests/golden/egc_stack_*.pt from the UNMODIFIED reference stack `experiments/mag/models.py::EGC`
(imported through oracle/shims; run in the build container):

    python -m oracle.make_golden_stack

Inputs, the full state_dict, the log-softmax output, the loss `nll_loss(out[train], y)` as the reference's training
loop computes it (experiments/mag/configs.py) and every gradient, in fp32 and fp64, in eval-free training mode with
dropout 0 (dropout draws are not reproducible across devices) - plus one eval-mode run with dropout 0.5 to pin that
`training=False` disables it."""
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
# name, N, E, hidden, layers, heads, bases, aggrs
CASES = [
    ("egc_stack_s", 150, 900, 64, 3, 8, 4, ["symnorm"]),
    ("egc_stack_m", 130, 800, 64, 3, 4, 4, ["symnorm", "max", "std"]),
]


def main():
    ref = rl.load()
    models = importlib.import_module("experiments.mag.models")
    for name, n, e, hidden, layers, heads, bases, aggrs in CASES:
        gen = torch.Generator().manual_seed(sum(ord(c) for c in name))
        ei = make_graph(n, e, 0, gen)
        perm = (ei[1] * n + ei[0]).argsort(stable=True)
        adj = ref.SparseTensor(row=ei[1][perm], col=ei[0][perm], sparse_sizes=(n, n), is_sorted=True)
        x = torch.randn(n, models.IN_FEATURES, generator=gen)
        y = torch.randint(0, models.OUT_TRUE, (n,), generator=gen)
        train_idx = torch.randperm(n, generator=gen)[: n // 2]
        torch.manual_seed(99)
        model = models.EGC(hidden, layers, 0.0, heads, bases, aggrs)
        rec = {"name": name, "n": n, "hidden": hidden, "layers": layers, "heads": heads, "bases": bases, "aggrs": aggrs,
               "edge_index": ei, "adj_rowptr": adj.csr()[0].clone(), "adj_col": adj.csr()[1].clone(), "x": x, "y": y,
               "train_idx": train_idx, "state_dict": {k: v.clone() for k, v in model.state_dict().items()}}
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            m = models.EGC(hidden, layers, 0.0, heads, bases, aggrs).to(dt)
            m.load_state_dict({k: v.to(dt) for k, v in rec["state_dict"].items()})
            m.train()
            xx = x.to(dt).requires_grad_(True)
            out = m(xx, adj)
            loss = F.nll_loss(out[train_idx], y[train_idx])
            params = list(m.named_parameters())
            grads = torch.autograd.grad(loss, [xx] + [p for _, p in params])
            rec[f"out_{tag}"], rec[f"loss_{tag}"], rec[f"grad_x_{tag}"] = out.detach(), loss.detach(), grads[0]
            for (pn, _), g in zip(params, grads[1:]):
                rec[f"grad_{pn}_{tag}"] = g
        m = models.EGC(hidden, layers, 0.5, heads, bases, aggrs)
        m.load_state_dict(rec["state_dict"])
        m.eval()
        with torch.no_grad():
            rec["out_eval_dropout_f32"] = m(x, adj)
        path = os.path.join(OUT_DIR, f"{name}.pt")
        torch.save(rec, path)
        print(f"{name:16s} loss {float(rec['loss_f64']):.6f}  {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
