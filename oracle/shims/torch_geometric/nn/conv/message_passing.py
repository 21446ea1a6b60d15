"""Minimal `MessagePassing` (PyG 2.0 behaviour the reference relies on).

* SparseTensor input + an overridden `message_and_aggregate` -> the fused call;
* otherwise `x_j = x.index_select(node_dim, edge_index[0])`, `index = edge_index[1]`,
  `dim_size = N`, then `message` -> `aggregate` -> `update` (identity);
* arguments are routed to the user hooks by parameter name, as PyG's inspector does.
"""
import inspect

import torch
from torch import Tensor
from torch_scatter import scatter
from torch_sparse import SparseTensor


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow: str = "source_to_target", node_dim: int = -2):
        super().__init__()
        self.aggr = aggr
        self.flow = flow
        self.node_dim = node_dim
        assert flow == "source_to_target", "oracle shim: only the default flow is restated"
        cls = type(self)
        self.fuse = cls.message_and_aggregate is not MessagePassing.message_and_aggregate

    @staticmethod
    def _params(fn, skip):
        return [p for p in list(inspect.signature(fn).parameters)[skip:]]

    def propagate(self, edge_index, size=None, **kwargs):
        if isinstance(edge_index, SparseTensor) and self.fuse:
            names = self._params(self.message_and_aggregate, 1)
            out = self.message_and_aggregate(edge_index, **{k: kwargs[k] for k in names if k in kwargs})
            return self.update(out)

        if isinstance(edge_index, SparseTensor):
            tgt, src, _ = edge_index.coo()
            n_tgt = edge_index.sparse_size(0)
        else:
            src, tgt = edge_index[0], edge_index[1]
            n_tgt = None

        collected = dict(kwargs)
        for name in self._params(self.message, 0):
            if name.endswith("_j") or name.endswith("_i"):
                data = kwargs[name[:-2]]
                if isinstance(data, Tensor):
                    dim = self.node_dim if self.node_dim >= 0 else data.dim() + self.node_dim
                    if n_tgt is None:
                        n_tgt = data.size(dim)
                    collected[name] = data.index_select(dim, src if name.endswith("_j") else tgt)
                else:
                    collected[name] = data
        if size is not None and size[1] is not None:
            n_tgt = size[1]
        collected.update(index=tgt, ptr=None, dim_size=n_tgt)

        msg = self.message(**{k: collected[k] for k in self._params(self.message, 0) if k in collected})
        agg_names = self._params(self.aggregate, 1)
        out = self.aggregate(msg, **{k: collected[k] for k in agg_names if k in collected})
        return self.update(out)

    def message(self, x_j: Tensor) -> Tensor:
        return x_j

    def aggregate(self, inputs: Tensor, index: Tensor, ptr=None, dim_size=None) -> Tensor:
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=self.aggr)

    def message_and_aggregate(self, adj_t, **kwargs):
        raise NotImplementedError

    def update(self, inputs: Tensor) -> Tensor:
        return inputs
