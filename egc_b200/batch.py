"""Mini-batch plumbing on the device: the two steps either side of the layer that the
reference gets from PyG on the CPU.

* `collate(graphs)` / `Batch` - block-diagonal collation of many small graphs, PyG `Batch.from_data_list` as driven
  by the reference's DataLoaders (/root/reference/experiments/zinc/configs.py:36-45,60-67,
  /root/reference/experiments/cifar/configs.py:42-53): node features concatenated, node ids of graph g shifted by the
  number of nodes before it, `batch[i]` = graph of node i, `ptr` = node offsets.
* `global_add_pool / global_mean_pool / global_max_pool(x, batch, size)` - the readout of
  /root/reference/experiments/zinc/models.py:46-53,73 as a deterministic segmented reduction with its own backward.

Both call libegc_b200 (csrc/batch.cu); there is no CPU path.
"""
from typing import List, Optional, Sequence, Tuple, Union

import torch
from torch import Tensor

from . import _lib
from ._lib import check, ptr
from .graph import _stream


class Batch:
    """What the layer stack needs of a PyG `Batch`: x, edge_index, batch, ptr, num_graphs."""

    def __init__(self, x: Optional[Tensor], edge_index: Tensor, batch: Tensor, node_ptr: Tensor, num_graphs: int):
        self.x, self.edge_index, self.batch, self.ptr, self.num_graphs = x, edge_index, batch, node_ptr, num_graphs

    @property
    def num_nodes(self) -> int:
        return int(self.batch.numel())


def _require_cuda(name: str, t: Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"egc_b200: `{name}` must be a CUDA tensor - this library has no CPU path")


def collate_arrays(edge_local: Tensor, edge_ptr: Tensor, node_ptr: Tensor, num_nodes: Optional[int] = None,
                   validate: bool = True) -> Tuple[Tensor, Tensor]:
    """(edge_index [2, E] int64 with global ids, batch [N] int64) from the concatenated graph-local edge lists.
    edge_local [2, E] int64; edge_ptr / node_ptr [G + 1] int32 exclusive prefix sums of per-graph edge / node counts.
    `num_nodes` (= node_ptr[-1], known to a caller that built the offsets on the host) and `validate=False` (skip the
    check that every local id lies inside its graph) each save one device-to-host read, i.e. one pipeline drain per
    batch; PyG's collation does not validate ids either."""
    lib = _lib.load()
    for n, t in (("edge_local", edge_local), ("edge_ptr", edge_ptr), ("node_ptr", node_ptr)):
        _require_cuda(n, t)
    if edge_local.dtype != torch.int64 or edge_local.dim() != 2 or edge_local.size(0) != 2:
        raise ValueError("edge_local must be an int64 tensor of shape [2, E]")
    if edge_ptr.dtype != torch.int32 or node_ptr.dtype != torch.int32 or edge_ptr.numel() != node_ptr.numel():
        raise ValueError("edge_ptr and node_ptr must be int32 tensors of the same length (num_graphs + 1)")
    g = int(node_ptr.numel()) - 1
    if g < 0:
        raise ValueError("node_ptr must hold at least one element")
    edge_local = edge_local.contiguous()
    e = int(edge_local.size(1))
    if num_nodes is None:
        ends = torch.stack([node_ptr[-1], edge_ptr[-1]]).cpu()       # one host read: N and the edge count check
        n = int(ends[0])
        if int(ends[1]) != e:
            raise ValueError(f"edge_ptr[-1] = {int(ends[1])} does not match the number of edges {e}")
    else:
        n = int(num_nodes)
    dev = edge_local.device
    out = torch.empty((2, e), dtype=torch.int64, device=dev)
    batch = torch.empty(n, dtype=torch.int64, device=dev)
    flags = torch.empty(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(lib.egc_collate_edges(ptr(edge_local[0]), ptr(edge_local[1]), ptr(edge_ptr.contiguous()),
                                    ptr(node_ptr.contiguous()), g, e, n, ptr(out[0]), ptr(out[1]), ptr(batch), ptr(flags),
                                    _stream()), "egc_collate_edges")
    if validate and int(flags.item()) & 1:
        raise ValueError("collate: an edge refers to a node id outside its graph")
    return out, batch


def collate(graphs: Sequence[Tuple[Optional[Tensor], Tensor, int]], device=None) -> Batch:
    """Batch.from_data_list for (x | None, edge_index [2, E_g] graph-local int64, num_nodes) triples.  The per-graph
    arrays are concatenated (one H2D copy each if they live on the host); the id shift and the batch vector are
    computed on the device."""
    if len(graphs) == 0:
        raise ValueError("collate needs at least one graph")
    device = torch.device(device) if device is not None else graphs[0][1].device
    counts = torch.tensor([[int(n), int(ei.size(1))] for _, ei, n in graphs], dtype=torch.int64)
    ptrs = torch.zeros((2, len(graphs) + 1), dtype=torch.int32)
    ptrs[:, 1:] = counts.cumsum(0).t().to(torch.int32)
    edge_local = torch.cat([ei for _, ei, _ in graphs], dim=1).to(device, non_blocking=True)
    ptrs = ptrs.to(device, non_blocking=True)
    x = None
    if graphs[0][0] is not None:
        x = torch.cat([xg for xg, _, _ in graphs], dim=0).to(device, non_blocking=True)
    edge_index, batch = collate_arrays(edge_local, ptrs[1], ptrs[0], num_nodes=int(counts[:, 0].sum()))
    return Batch(x, edge_index, batch, ptrs[0], len(graphs))


def segment_ptr(batch: Tensor, num_graphs: int) -> Tensor:
    """int32 [num_graphs + 1] node offsets of a sorted PyG batch vector (raises if it is not sorted)."""
    lib = _lib.load()
    _require_cuda("batch", batch)
    if batch.dtype != torch.int64 or batch.dim() != 1:
        raise ValueError("batch must be a 1-D int64 tensor")
    out = torch.empty(num_graphs + 1, dtype=torch.int32, device=batch.device)
    flags = torch.empty(1, dtype=torch.int32, device=batch.device)
    with torch.cuda.device(batch.device):
        check(lib.egc_segment_ptr(ptr(batch.contiguous()), int(batch.numel()), int(num_graphs), ptr(out), ptr(flags),
                                  _stream()), "egc_segment_ptr")
    f = int(flags.item())
    if f & 2:
        raise ValueError("batch holds a graph id outside [0, num_graphs)")
    if f & 1:
        raise ValueError("batch must be sorted (nodes of one graph contiguous), as PyG's collation produces it")
    return out


class _SegmentPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, node_ptr, mode):
        lib = _lib.load()
        _require_cuda("x", x)
        if x.dtype != torch.float32 or x.dim() != 2:
            raise TypeError("global pooling expects a float32 [num_nodes, channels] tensor")
        x = x.contiguous()
        g, f = int(node_ptr.numel()) - 1, int(x.size(1))
        out = torch.empty((g, f), dtype=torch.float32, device=x.device)
        arg = torch.empty((g, f), dtype=torch.int32, device=x.device) if mode == _lib.POOL_CODES["max"] else None
        with torch.cuda.device(x.device):
            check(lib.egc_segment_pool_fwd(ptr(x), ptr(node_ptr), g, f, mode, ptr(out), ptr(arg), _stream()),
                  "egc_segment_pool_fwd")
        ctx.save_for_backward(node_ptr, arg)
        ctx.mode, ctx.n = mode, int(x.size(0))
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        node_ptr, arg = ctx.saved_tensors
        d_out = d_out.contiguous()
        g, f = int(d_out.size(0)), int(d_out.size(1))
        d_x = torch.zeros((ctx.n, f), dtype=torch.float32, device=d_out.device)     # nodes outside every segment: 0
        with torch.cuda.device(d_out.device):
            check(lib.egc_segment_pool_bwd(ptr(d_out), ptr(node_ptr), ptr(arg), g, f, ctx.mode, ptr(d_x), _stream()),
                  "egc_segment_pool_bwd")
        return d_x, None, None


def _pool(x: Tensor, batch: Union[Tensor, Batch], size: Optional[int], mode: str) -> Tensor:
    if isinstance(batch, Batch):
        node_ptr = batch.ptr
    elif batch.dtype == torch.int32:                      # already node offsets [G + 1]
        node_ptr = batch
    else:
        if size is None:                                  # PyG: size = batch.max() + 1 (one host read)
            size = int(batch[-1].item()) + 1 if batch.numel() else 0
        node_ptr = segment_ptr(batch, size)
    return _SegmentPool.apply(x, node_ptr, _lib.POOL_CODES[mode])


def global_add_pool(x: Tensor, batch, size: Optional[int] = None) -> Tensor:
    return _pool(x, batch, size, "sum")


def global_mean_pool(x: Tensor, batch, size: Optional[int] = None) -> Tensor:
    return _pool(x, batch, size, "mean")


def global_max_pool(x: Tensor, batch, size: Optional[int] = None) -> Tensor:
    return _pool(x, batch, size, "max")


# ---------------------------------------------------------------------------------------------------
# fixed-shape batches: what lets a whole training step (collation, CSR / CSC build, layers, readout) replay from ONE
# CUDA graph.  The host side (the DataLoader's collate function) pads every batch to the same node and edge counts with
# one extra "pad graph" that no readout segment covers: its nodes only talk to each other, so nothing computed for the
# real graphs changes, and every gradient they could contribute is multiplied by a zero grad_out.
# ---------------------------------------------------------------------------------------------------
def pad_batch(x: Tensor, edge_local: Tensor, ptrs: Tensor, n_cap: int, e_cap: int):
    """HOST tensors of one collated batch -> the same batch padded to exactly `n_cap` nodes and `e_cap` edges.

    x [n, F]; edge_local [2, E] graph-local ids; ptrs int32 [2, G + 1] (row 0: node offsets, row 1: edge offsets).
    Returns (x_pad [n_cap, F], edge_pad [2, e_cap], ptrs_pad [2, G + 2], nnz_expected): graph G is the pad graph - a ring
    over its n_cap - n nodes carrying the e_cap - E extra edges.  `ptrs_pad[0, :G + 1]` are the readout segments.
    `nnz_expected` = edges that are not self-loops + one loop per node = nnz of the prepared graph (add_self_loops)."""
    n, e = int(x.size(0)), int(edge_local.size(1))
    p, q = n_cap - n, e_cap - e
    if p < 2 or q < 1:
        raise ValueError(f"pad_batch: capacity too small (nodes {n} -> {n_cap}, edges {e} -> {e_cap}; need >= 2 pad nodes "
                         "and >= 1 pad edge)")
    if q > p * (_lib.EGC_CHUNK_EDGES // 2):
        raise ValueError(f"pad_batch: {q} pad edges over {p} pad nodes would create rows longer than the chunk size")
    x_pad = torch.zeros((n_cap, x.size(1)), dtype=x.dtype)
    x_pad[:n] = x
    k = torch.arange(q, dtype=torch.int64)
    ring = torch.stack([(k + p - 1) % p, k % p])                 # edge 0 leaves the LAST pad node: the largest id is always used
    edge_pad = torch.cat([edge_local.to(torch.int64), ring], 1)
    g = ptrs.size(1) - 1
    ptrs_pad = torch.empty((2, g + 2), dtype=torch.int32)
    ptrs_pad[:, :g + 1] = ptrs
    ptrs_pad[0, g + 1], ptrs_pad[1, g + 1] = n_cap, e_cap
    loops = int((edge_local[0] == edge_local[1]).sum())          # graph-local ids: a loop is a loop
    return x_pad, edge_pad, ptrs_pad, (e - loops) + q + n_cap
