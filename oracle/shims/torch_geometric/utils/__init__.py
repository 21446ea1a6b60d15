"""`torch_geometric.utils` stand-ins."""
from typing import Optional, Tuple

import torch
from torch import Tensor


def add_remaining_self_loops(edge_index: Tensor, edge_weight: Optional[Tensor] = None,
                             fill_value: float = 1.0, num_nodes: Optional[int] = None
                             ) -> Tuple[Tensor, Optional[Tensor]]:
    """Drop existing self-loops, append exactly one loop per node `< N` at the end.

    `N = num_nodes` or `edge_index.max() + 1`.  A pre-existing loop's weight is kept for
    that node (last one wins); all other loops get `fill_value`.
    """
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    row, col = edge_index[0], edge_index[1]
    off_diag = row != col
    loop_index = torch.arange(num_nodes, dtype=row.dtype, device=row.device)
    loop_index = loop_index.unsqueeze(0).repeat(2, 1)
    if edge_weight is not None:
        loop_weight = edge_weight.new_full((num_nodes,), fill_value)
        on_diag = ~off_diag
        loop_weight[row[on_diag]] = edge_weight[on_diag]
        edge_weight = torch.cat([edge_weight[off_diag], loop_weight], dim=0)
    edge_index = torch.cat([edge_index[:, off_diag], loop_index], dim=1)
    return edge_index, edge_weight


def to_undirected(edge_index: Tensor, num_nodes: Optional[int] = None) -> Tensor:
    """Symmetrise and coalesce (sorted by (row, col), duplicates removed)."""
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    row = torch.cat([edge_index[0], edge_index[1]])
    col = torch.cat([edge_index[1], edge_index[0]])
    key = torch.unique(row * num_nodes + col, sorted=True)
    row = torch.div(key, num_nodes, rounding_mode="floor")
    return torch.stack([row, key - row * num_nodes])
