"""Shared helpers for the test-suite (graph generators, fixtures, comparisons)."""
import glob
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases(prefix=None):
    """EGConv fixtures by default; `prefix="paper_"` selects the EfficientGraphConv (paper variant) fixtures."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))
    if prefix is None:
        return [n for n in names if not n.startswith(("paper_", "regconv_", "egc_stack_", "arxivnet_"))]
    return [n for n in names if n.startswith(prefix)]


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


def random_graph(n, e, seed, hub=0, isolated=2, self_loops=5, dups=20):
    """edge_index [2, E'] with a few self-loops, duplicated edges, trailing isolated nodes, optional hubs."""
    gen = torch.Generator().manual_seed(seed)
    hi = max(n - isolated, 1)
    src = torch.randint(0, hi, (e,), generator=gen)
    dst = torch.randint(0, hi, (e,), generator=gen)
    src[:self_loops] = dst[:self_loops]
    src = torch.cat([src, src[self_loops:self_loops + dups]])
    dst = torch.cat([dst, dst[self_loops:self_loops + dups]])
    if hub:
        others = torch.randperm(hi, generator=gen)[:hub]
        src = torch.cat([src, others, torch.full((others.numel(),), 3)])
        dst = torch.cat([dst, torch.full((others.numel(),), 7), others])
    return torch.stack([src, dst])


def to_adj_csr(edge_index, n, value=None):
    """(rowptr, col, value) of adj_t: rows = targets, sorted by (target, source) - experiments/utils.py:93."""
    perm = (edge_index[1] * n + edge_index[0]).argsort(stable=True)
    src, dst = edge_index[0][perm], edge_index[1][perm]
    rowptr = torch.zeros(n + 1, dtype=torch.long)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n), 0)
    return rowptr, src, (value[perm] if value is not None else None)


def rel_err(a, b):
    """max |a-b| relative to max |b| (the 1e-5 bar of BASELINE.json is on this quantity)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = b.abs().max().clamp(min=1e-30)
    return ((a - b).abs().max() / denom).item()


def restrict_rows(g, keep_rows):
    """Oracle graph with the entries of the target rows in `keep_rows` (bool [n_dst]) only; every other row becomes
    empty.  Weights (symnorm / linear) keep their GLOBAL values, so the kept rows aggregate exactly as in `g`."""
    from oracle.restatement import OracleGraph
    row = g.row
    sel = keep_rows[row]
    cnt = torch.where(keep_rows, g.rowptr[1:] - g.rowptr[:-1], torch.zeros_like(g.rowptr[1:]))
    rowptr = torch.cat([cnt.new_zeros(1), cnt.cumsum(0)])
    pick = lambda t: t[sel] if t is not None else None  # noqa: E731
    return OracleGraph(rowptr, g.col[sel], g.n_dst, g.n_src, val_sym=pick(g.val_sym), val_lin=pick(g.val_lin),
                       deg=g.deg, dis=g.dis)
