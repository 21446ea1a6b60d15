"""TEST INFRASTRUCTURE - CPU restatement of the reference's heterogeneous layer `REGConv`
(/root/reference/experiments/rmag/models.py:75-148) on top of oracle/restatement.py.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this.

Pinned against the UNMODIFIED reference class executed through oracle/shims (tests/test_hetero.py, where
/root/reference is mounted) and against tests/golden/regconv_*.pt generated from it by oracle/make_golden.py.
"""
from typing import Dict, Sequence, Tuple

import torch
from torch import Tensor
from torch.nn import Linear, ModuleDict, Parameter

from . import restatement as R

NODE_TYPES = ("author", "field_of_study", "institution", "paper")                 # ref rmag/models.py:17
EDGE_TYPES = (("author", "affiliated_with", "institution"), ("institution", "to", "author"),
              ("author", "writes", "paper"), ("paper", "to", "author"), ("paper", "cites", "paper"),
              ("paper", "has_topic", "field_of_study"), ("field_of_study", "to", "paper"))   # ref :18-26


class REGConvOracle(torch.nn.Module):
    def __init__(self, in_channels, out_channels, num_heads, num_bases, node_types: Sequence[str] = NODE_TYPES,
                 edge_types: Sequence[Tuple[str, str, str]] = EDGE_TYPES):
        super().__init__()
        self.in_channels, self.out_channels, self.num_heads, self.num_bases = in_channels, out_channels, num_heads, num_bases
        self.bases_weight = Parameter(torch.empty(in_channels, (out_channels // num_heads) * num_bases))   # ref :84-86
        self.rel_combs = ModuleDict({f"{k[0]}_{k[1]}_{k[2]}": Linear(in_channels, 2 * num_heads * num_bases)
                                     for k in edge_types})                                               # ref :89-97
        self.root_combs = ModuleDict({k: Linear(in_channels, num_heads * num_bases) for k in node_types})   # ref :99-101
        a = (6.0 / (self.bases_weight.size(-2) + self.bases_weight.size(-1))) ** 0.5
        with torch.no_grad():
            self.bases_weight.uniform_(-a, a)                                                            # glorot, ref :106

    def forward(self, x_dict: Dict[str, Tensor], csr_dict: Dict[Tuple[str, str, str], Tuple[Tensor, Tensor, int]]):
        """csr_dict[(src, rel, dst)] = (rowptr [N_dst + 1], col [E] source ids, n_src) of adj_t (rows = targets)."""
        h = self.num_heads
        bases = {t: x @ self.bases_weight for t, x in x_dict.items()}                                    # ref :116-118
        out = {}
        for t, x in x_dict.items():                                                                      # ref :120-132
            out[t] = R.combine(self.root_combs[t](x), bases[t].unsqueeze(1), None, h)
        for key, (rowptr, col, n_src) in csr_dict.items():                                               # ref :134-146
            g = R.graph_from_csr(rowptr, col, None, n_src, False, False, True, dtype=bases[key[0]].dtype)
            agg, _ = R.aggregate(g, bases[key[0]], ["mean", "max"])
            w = self.rel_combs[f"{key[0]}_{key[1]}_{key[2]}"](x_dict[key[2]])
            out[key[2]] = out[key[2]] + R.combine(w, agg, None, h)
        return out


def random_hetero_graph(sizes: Dict[str, int], edges_per_relation: int, seed: int, edge_types=EDGE_TYPES):
    """{(src, rel, dst): (rowptr, col, n_src)} with rows sorted by (target, source), duplicates kept, some empty rows."""
    gen = torch.Generator().manual_seed(seed)
    out = {}
    for (s, r, d) in edge_types:
        n_dst, n_src = sizes[d], sizes[s]
        dst = torch.randint(0, max(n_dst - 2, 1), (edges_per_relation,), generator=gen)     # last rows stay empty
        src = torch.randint(0, n_src, (edges_per_relation,), generator=gen)
        order = torch.argsort(dst * n_src + src, stable=True)
        dst, src = dst[order], src[order]
        rowptr = torch.zeros(n_dst + 1, dtype=torch.long)
        rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_dst), 0)
        out[(s, r, d)] = (rowptr, src, n_src)
    return out
