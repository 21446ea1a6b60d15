"""TEST INFRASTRUCTURE - CPU restatement of the mini-batch steps either side of the layer.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.

The arithmetic lives in an un-vendored third party: torch-geometric==2.0 (pinned by /root/reference/Dockerfile:53-54)
- `Batch.from_data_list` behind the reference's DataLoaders (/root/reference/experiments/zinc/configs.py:36-45,60-67,
/root/reference/experiments/cifar/configs.py:42-53) and `global_add_pool / global_mean_pool / global_max_pool`
(/root/reference/experiments/zinc/models.py:3-8,46-53,73).  Published semantics restated here:
  * collation: x concatenated; `edge_index` of graph g shifted by the number of nodes of graphs 0..g-1; `batch[i]` = g;
    `ptr` = cumulative node counts;
  * pooling = torch_scatter.scatter(x, batch, dim=0, dim_size=size, reduce): sum; mean = sum / max(count, 1);
    max with empty graphs -> 0 and the gradient sent to the first maximal element.
Parity unpinned by the reference itself (it has no tests for these steps); pinned here by hand-computed cases in
tests/test_oracle.py.
"""
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor


def collate(graphs: Sequence[Tuple[Optional[Tensor], Tensor, int]]):
    """(x | None, edge_index [2, E] int64, batch [N] int64, ptr [G + 1] int64) - plain Python loop over the graphs."""
    xs, eis, batch, ptr = [], [], [], [0]
    for g, (x, ei, n) in enumerate(graphs):
        if ei.numel() and (int(ei.min()) < 0 or int(ei.max()) >= n):
            raise ValueError("collate: an edge refers to a node id outside its graph")
        eis.append(ei + ptr[-1])
        batch.append(torch.full((n,), g, dtype=torch.int64))
        if x is not None:
            xs.append(x)
        ptr.append(ptr[-1] + n)
    return (torch.cat(xs, 0) if xs else None, torch.cat(eis, 1), torch.cat(batch), torch.tensor(ptr, dtype=torch.int64))


class _FirstMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, batch, size):
        out = torch.zeros((size, x.size(1)), dtype=x.dtype)
        arg = torch.full((size, x.size(1)), -1, dtype=torch.int64)
        for i in range(x.size(0)):                      # strict > keeps the first maximal element
            g = int(batch[i])
            better = (arg[g] < 0) | (x[i] > out[g])
            out[g] = torch.where(better, x[i], out[g])
            arg[g] = torch.where(better, torch.full_like(arg[g], i), arg[g])
        ctx.save_for_backward(arg)
        ctx.n = x.size(0)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (arg,) = ctx.saved_tensors
        d_x = torch.zeros((ctx.n, d_out.size(1)), dtype=d_out.dtype)
        cols = torch.arange(d_out.size(1)).expand_as(arg)
        ok = arg >= 0
        d_x[arg[ok], cols[ok]] = d_out[ok]
        return d_x, None, None


def global_pool(x: Tensor, batch: Tensor, size: Optional[int], reduce: str) -> Tensor:
    if size is None:
        size = int(batch.max()) + 1 if batch.numel() else 0
    if reduce in ("sum", "add", "mean"):
        out = torch.zeros((size, x.size(1)), dtype=x.dtype).index_add(0, batch, x)
        if reduce == "mean":
            cnt = torch.bincount(batch, minlength=size).clamp(min=1).to(x.dtype)
            out = out / cnt[:, None]
        return out
    if reduce == "max":
        return _FirstMax.apply(x, batch, size)
    raise ValueError(reduce)


# ---------------------------------------------------------------------------------------------
# synthetic mini-batches of the shapes named by BASELINE.json configs 1 and 5
# ---------------------------------------------------------------------------------------------
def zinc_like_graphs(num_graphs: int = 128, seed: int = 0) -> List[Tuple[Tensor, Tensor, int]]:
    """Molecule-like graphs: ~N(23.2, 4.5) nodes clipped to [9, 37], a random tree plus a few ring closures (mean degree
    ~2.15), both directions; x = integer atom type in [0, 28) (ref zinc/models.py:13,27)."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num_graphs):
        n = int(torch.clamp(torch.round(torch.randn((), generator=gen) * 4.5 + 23.2), 9, 37))
        parent = torch.tensor([int(torch.randint(0, i, (), generator=gen)) for i in range(1, n)], dtype=torch.int64)
        child = torch.arange(1, n, dtype=torch.int64)
        n_ring = max(1, int(round(0.075 * n)))
        a = torch.randint(0, n, (n_ring,), generator=gen)
        b = (a + torch.randint(2, max(n - 1, 3), (n_ring,), generator=gen)) % n
        keep = a != b
        src = torch.cat([parent, a[keep]])
        dst = torch.cat([child, b[keep]])
        ei = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
        out.append((torch.randint(0, 28, (n, 1), generator=gen), ei, n))
    return out


def cifar_like_graphs(num_graphs: int = 128, seed: int = 0, k: int = 8) -> List[Tuple[Tensor, Tensor, int]]:
    """Superpixel-like graphs: U{85..150} nodes at uniform 2-D positions, k = 8 nearest neighbours as in-edges of every
    node (ref cifar/configs.py:42-53), x = [rgb(3) | pos(2)]."""
    gen = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(num_graphs):
        n = int(torch.randint(85, 151, (), generator=gen))
        pos = torch.rand((n, 2), generator=gen)
        d = torch.cdist(pos, pos)
        d.fill_diagonal_(float("inf"))
        nbr = d.topk(k, largest=False).indices                      # [n, k] sources of node i
        ei = torch.stack([nbr.reshape(-1), torch.arange(n).repeat_interleave(k)])
        out.append((torch.cat([torch.rand((n, 3), generator=gen), pos], 1), ei, n))
    return out
