"""Generate tests/golden/regconv_*.pt from the UNMODIFIED reference `REGConv`
(/root/reference/experiments/rmag/models.py:75-148, imported through oracle/shims; run in the build container).

    python -m oracle.make_golden_hetero

Stored per case: node-type sizes, the relations' CSRs, x per node type, grad_out per node type, the state_dict, and the
reference's outputs and gradients (x per type + every parameter) in fp32 and fp64.
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hetero as OH  # noqa: E402
from oracle import reference_loader as rl  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")
# name, sizes, edges per relation, F_in, F_out, H, B
CASES = [
    ("regconv_small", {"author": 70, "field_of_study": 24, "institution": 9, "paper": 48}, 260, 32, 64, 4, 4),
    ("regconv_mag_shape", {"author": 90, "field_of_study": 30, "institution": 10, "paper": 60}, 500, 128, 128, 8, 4),
]


def run(dtype, name, sizes, e, f_in, f_out, h, b, ref_mod, ST, state=None):
    gen = torch.Generator().manual_seed(sum(ord(c) for c in name))
    csr = OH.random_hetero_graph(sizes, e, seed=sum(ord(c) for c in name) + 1)
    x = {t: torch.randn(n, f_in, generator=gen).to(dtype).requires_grad_(True) for t, n in sizes.items()}
    go = {t: torch.randn(n, f_out, generator=gen).to(dtype) for t, n in sizes.items()}
    torch.manual_seed(4321)
    conv = ref_mod.REGConv(f_in, f_out, h, b)
    if state is not None:
        conv.load_state_dict(state)
    conv = conv.to(dtype)
    adj = {k: ST(rowptr=rp, col=col, sparse_sizes=(rp.numel() - 1, n_src), is_sorted=True) for k, (rp, col, n_src) in csr.items()}
    out = conv(x, adj)
    params = list(conv.named_parameters())
    loss = sum((out[t] * go[t]).sum() for t in sizes)
    grads = torch.autograd.grad(loss, [x[t] for t in sizes] + [p for _, p in params])
    rec = {"out": {t: out[t].detach() for t in sizes},
           "grad_x": {t: g for t, g in zip(sizes, grads[:len(sizes)])},
           "grad_p": {n_: g for (n_, _), g in zip(params, grads[len(sizes):])}}
    return rec, csr, x, go, {k: v.detach().float() for k, v in conv.state_dict().items()}


def main():
    rl.load()
    ref_mod = importlib.import_module("experiments.rmag.models")
    ST = rl.shims().SparseTensor
    os.makedirs(OUT_DIR, exist_ok=True)
    for name, sizes, e, f_in, f_out, h, b in CASES:
        r32, csr, x, go, state = run(torch.float32, name, sizes, e, f_in, f_out, h, b, ref_mod, ST)
        r64, *_ = run(torch.float64, name, sizes, e, f_in, f_out, h, b, ref_mod, ST, state)
        rec = {"sizes": sizes, "f_in": f_in, "f_out": f_out, "heads": h, "bases": b,
               "csr": {k: (rp, col, n_src) for k, (rp, col, n_src) in csr.items()},
               "x": {t: v.detach() for t, v in x.items()}, "grad_out": go, "state_dict": state, "f32": r32, "f64": r64}
        path = os.path.join(OUT_DIR, name + ".pt")
        torch.save(rec, path)
        print(name, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
