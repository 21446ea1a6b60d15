"""CPU restatement of the reference EGConv hot path (TEST INFRASTRUCTURE ONLY).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import
this module; nothing under `egc_b200/` does, and the product path has no CPU fallback.

It restates, in CSR form and plain torch (CPU, fp32 or fp64), what
`/root/reference/experiments/optimized_layers.py:124-278` computes through
torch_geometric / torch_scatter / torch_sparse, so that the CUDA kernels can be
checked stage by stage on the GPU box where `/root/reference` does not exist:

    graph preparation  -> `graph_from_edge_index`, `graph_from_csr`     (ref :127-175)
    projections        -> `project`                                     (ref :180-184)
    aggregation        -> `aggregate`                                   (ref :215-278)
    combination + bias -> `combine`                                     (ref :195-210)
    whole layer        -> `egconv_forward`, class `EGConvOracle`        (ref :74-210)
    analytic backward  -> `egconv_backward`  (autograd of the above, written out)

PARITY PIN: the reference has no tests, golden vectors or fixtures for this layer
 - "parity unpinned" by the reference's own tests.  The pin used instead:
this restatement is asserted equal to the reference source itself, executed unmodified
through `oracle/shims` by `oracle/reference_loader.py` (tests/test_oracle.py, runs
wherever /root/reference is present), and against `tests/golden/*.pt`, which
`oracle/make_golden.py` generated from that same unmodified reference source.
"""
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

AGGREGATORS = ("sum", "mean", "symnorm", "min", "max", "var", "std")   # ref :93
STD_EPS = 1e-5                                                          # ref :244,273


# ----------------------------------------------------------------------------------------
# graph preparation
# ----------------------------------------------------------------------------------------
@dataclass
class OracleGraph:
    """Target-major CSR of the graph the layer aggregates over (after self-loop handling).

    rowptr [N_dst+1], col [E] (source ids, in the order the reference would visit them),
    val_sym [E] symmetric-normalisation weight per nnz (None unless requested),
    val_lin [E] weight applied by the non-symnorm aggregators (None = unweighted),
    deg [N_dst] the degree `gcn_norm` used, dis = deg^-1/2 (inf -> 0).
    """
    rowptr: Tensor
    col: Tensor
    n_dst: int
    n_src: int
    val_sym: Optional[Tensor] = None
    val_lin: Optional[Tensor] = None
    deg: Optional[Tensor] = None
    dis: Optional[Tensor] = None

    @property
    def row(self) -> Tensor:
        cnt = self.rowptr[1:] - self.rowptr[:-1]
        return torch.repeat_interleave(torch.arange(self.n_dst), cnt)

    @property
    def nnz(self) -> int:
        return int(self.col.numel())


def _rowptr_from_sorted_rows(row: Tensor, n: int) -> Tensor:
    rp = torch.zeros(n + 1, dtype=torch.long)
    if row.numel():
        rp[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    return rp


def _inv_sqrt_degree(deg: Tensor) -> Tensor:
    # ref: gcn_norm -> deg.pow(-0.5), inf -> 0  
    dis = deg.pow(-0.5)
    return dis.masked_fill(dis == float("inf"), 0.0)


def graph_from_edge_index(edge_index: Tensor, num_nodes: int, symnorm: bool, add_self_loops: bool,
                          dtype=torch.float32) -> OracleGraph:
    """`edge_index` Tensor input (ref :128-141 symnorm branch, :159-166 self-loop branch).

    symnorm: loops for every node `< num_nodes` (gcn_norm is given num_nodes, ref :134).
    otherwise: `add_remaining_self_loops(edge_index)` WITHOUT num_nodes (ref :164), so only
    nodes `<= edge_index.max()` get a loop.  Existing loops are collapsed to one; the
    appended loops come last in edge order, so a stable sort by target keeps the
    reference's visiting order inside every row (first-wins ties of min/max).
    """
    src, dst = edge_index[0].long(), edge_index[1].long()
    if add_self_loops:
        n_loops = num_nodes if symnorm else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
        keep = src != dst
        loops = torch.arange(n_loops)
        src = torch.cat([src[keep], loops])
        dst = torch.cat([dst[keep], loops])
    order = torch.argsort(dst, stable=True)
    src, dst = src[order], dst[order]
    g = OracleGraph(_rowptr_from_sorted_rows(dst, num_nodes), src, num_nodes, num_nodes)
    if symnorm:
        g.deg = torch.zeros(num_nodes, dtype=dtype).index_add_(0, dst, torch.ones(dst.numel(), dtype=dtype))
        g.dis = _inv_sqrt_degree(g.deg)
        g.val_sym = g.dis[src] * torch.ones(dst.numel(), dtype=dtype) * g.dis[dst]   # dis[row]*w*dis[col]
    return g


def graph_from_csr(rowptr: Tensor, col: Tensor, value: Optional[Tensor], n_src: int, symnorm: bool,
                   add_self_loops: bool, multi_aggr: bool, dtype=torch.float32) -> OracleGraph:
    """SparseTensor (`adj_t`, rows = targets) input (ref :143-156 symnorm, :168-175 fill_diag).

    fill_diag removes every diagonal entry and inserts one per `i < min(rows, cols)` in
    sorted position.  gcn_norm: value (ones if absent) -> deg = row sums ->
    `(value * dis[row]) * dis[col]`.  The non-symnorm aggregators see the un-normalised
    matrix WITHOUT values when symnorm is also requested (`set_value(None)`, ref :253-254);
    if symnorm is the only aggregator nothing else is aggregated; if symnorm is not
    requested the input's own values (if any) weight every aggregator (ref :256-258).
    """
    n_dst = rowptr.numel() - 1
    rowptr, col = rowptr.long(), col.long()
    row = torch.repeat_interleave(torch.arange(n_dst), rowptr[1:] - rowptr[:-1])
    val = value
    if add_self_loops:
        keep = row != col
        d = torch.arange(min(n_dst, n_src))
        row2 = torch.cat([row[keep], d])
        col2 = torch.cat([col[keep], d])
        order = torch.argsort(row2 * n_src + col2, stable=True)
        row, col = row2[order], col2[order]
        if symnorm or val is not None:
            base = val[keep] if val is not None else torch.ones(int(keep.sum()), dtype=dtype)
            val = torch.cat([base, torch.ones(d.numel(), dtype=base.dtype)])[order]
    g = OracleGraph(_rowptr_from_sorted_rows(row, n_dst), col, n_dst, n_src)
    if symnorm:
        v = val if val is not None else torch.ones(col.numel(), dtype=dtype)
        g.deg = torch.zeros(n_dst, dtype=v.dtype).index_add_(0, row, v)
        g.dis = _inv_sqrt_degree(g.deg)
        g.val_sym = (v * g.dis[row]) * g.dis[col]
        g.val_lin = None
    else:
        g.val_lin = val if value is not None else None
    return g


# ----------------------------------------------------------------------------------------
# segment reductions over CSR rows
# ----------------------------------------------------------------------------------------
class _SegmentExtremum(torch.autograd.Function):
    """Row-wise max/min of `msgs[E, F]`; empty row -> 0; first nnz wins ties; the gradient
    goes to that one nnz only (torch_scatter / torch_sparse semantics)."""

    @staticmethod
    def forward(ctx, msgs: Tensor, row: Tensor, n: int, is_max: bool):
        e, f = msgs.shape
        idx = row.view(-1, 1).expand(e, f)
        ext = msgs.new_zeros((n, f)).scatter_reduce(0, idx, msgs, "amax" if is_max else "amin", include_self=False)
        pos = torch.arange(e).view(-1, 1).expand(e, f)
        first = torch.full((n, f), e, dtype=torch.long).scatter_reduce(
            0, idx, torch.where(msgs == ext[row], pos, e), "amin", include_self=True)
        out = torch.where(first < e, ext, torch.zeros_like(ext))
        ctx.save_for_backward(first)
        ctx.e = e
        ctx.mark_non_differentiable(first)
        return out, first

    @staticmethod
    def backward(ctx, g_out, _):
        (first,) = ctx.saved_tensors
        g = g_out.new_zeros((ctx.e + 1, first.shape[1]))
        g.scatter_(0, first, g_out)
        return g[: ctx.e], None, None, None


def aggregate(g: OracleGraph, bases: Tensor, aggrs: Sequence[str]) -> Tuple[Tensor, dict]:
    """All requested aggregators over every target row: returns `[N_dst, A, B*D]` (ref :215-278)
    and `{"max": argpos, "min": argpos}` (nnz position of the winning element, E = none)."""
    row, n = g.row, g.n_dst
    xj = bases.index_select(0, g.col)                       # PyG lift / spmm gather (ref :191)
    lin = xj if g.val_lin is None else xj * g.val_lin.view(-1, 1).to(xj.dtype)
    cnt = (g.rowptr[1:] - g.rowptr[:-1]).clamp(min=1).to(bases.dtype).view(-1, 1)
    zeros = bases.new_zeros((n, bases.shape[1]))
    out: List[Tensor] = []
    args = {}
    for a in aggrs:
        if a == "sum":                                      # ref :224-225 / :276
            r = zeros.index_add(0, row, lin)
        elif a == "symnorm":                                # ref :226-230 / :261-263
            assert g.val_sym is not None
            r = zeros.index_add(0, row, xj * g.val_sym.view(-1, 1).to(xj.dtype))
        elif a == "mean":                                   # ref :231-232; divides by nnz count
            r = zeros.index_add(0, row, lin) / cnt
        elif a in ("min", "max"):                           # ref :233-236
            r, args[a] = _SegmentExtremum.apply(lin, row, n, a == "max")
        elif a in ("var", "std"):                           # ref :237-244 / :268-274
            mean = zeros.index_add(0, row, lin) / cnt
            sq = xj * xj
            if g.val_lin is not None:
                sq = sq * g.val_lin.view(-1, 1).to(xj.dtype)
            mean_sq = zeros.index_add(0, row, sq) / cnt
            r = mean_sq - mean * mean
            if a == "std":
                r = torch.sqrt(torch.relu(r) + STD_EPS)
        else:
            raise ValueError(f'Unknown aggregator "{a}".')   # ref :246
        out.append(r)
    return torch.stack(out, dim=1), args                    # ref :249 / :278


# ----------------------------------------------------------------------------------------
# dense stages
# ----------------------------------------------------------------------------------------
def project(x: Tensor, bases_weight: Tensor, comb_weight: Tensor, comb_bias: Optional[Tensor],
            sigmoid: bool = False) -> Tuple[Tensor, Tensor]:
    """bases = x @ W_b (ref :180); weightings = Linear(x) (ref :182), optional sigmoid (:183-184)."""
    bases = x @ bases_weight
    w = torch.nn.functional.linear(x, comb_weight, comb_bias)
    return bases, (torch.sigmoid(w) if sigmoid else w)


def combine(weightings: Tensor, aggregated: Tensor, bias: Optional[Tensor], num_heads: int) -> Tensor:
    """out[n,h,:] = sum_{a,b} w[n,h,a*B+b] * agg[n,a,b*D:(b+1)*D]; head-major concat; + bias (ref :195-208)."""
    n, a, bd = aggregated.shape
    hab = weightings.shape[1]
    ab = hab // num_heads
    d = a * bd // ab
    out = torch.matmul(weightings.view(n, num_heads, ab), aggregated.reshape(n, ab, d)).reshape(n, num_heads * d)
    return out + bias if bias is not None else out


def egconv_forward(x, g: OracleGraph, bases_weight, comb_weight, comb_bias, bias, aggrs, num_heads,
                   sigmoid=False) -> Tensor:
    bases, w = project(x, bases_weight, comb_weight, comb_bias, sigmoid)
    agg, _ = aggregate(g, bases, aggrs)
    return combine(w, agg, bias, num_heads)


# ----------------------------------------------------------------------------------------
# analytic backward (what autograd does to the above, written out per stage)
# ----------------------------------------------------------------------------------------
def egconv_backward(x, g: OracleGraph, bases_weight, comb_weight, comb_bias, bias, aggrs, num_heads,
                    grad_out: Tensor, sigmoid=False) -> dict:
    """Closed-form gradients of `egconv_forward`; the decomposition the CUDA backward uses.

    Per target i (CSR pass):  d_w = g . agg^T,  d_agg = w^T . g, then per aggregator the
    gradient w.r.t. each incoming message, expressed as target-side vectors
        t_sym[i] (symnorm), t_lin[i] (sum, mean, the -2*mean term of var/std),
        t_sq[i]  (the x_j^2 term of var/std)
    plus single-element routing for min/max.  Per source j (CSC pass):
        d_bases[j] = sum_e val_sym[e] t_sym[i_e] + val_lin[e] (t_lin[i_e] + 2 x_j t_sq[i_e]) + routed.
    Dense: d_x = d_bases W_b^T + d_lin W_c, dW_b = x^T d_bases, dW_c = d_lin^T x, db_c = sum d_lin,
    dbias = sum grad_out.  (autograd of ref :180-208; torch_sparse spmm backward, App. A-6.)
    """
    n = g.n_dst
    bases, w = project(x, bases_weight, comb_weight, comb_bias, sigmoid)
    agg, args = aggregate(g, bases, aggrs)
    a_n, bd = agg.shape[1], agg.shape[2]
    ab = w.shape[1] // num_heads
    d = a_n * bd // ab
    g3 = grad_out.view(n, num_heads, d)
    d_w = torch.matmul(g3, agg.reshape(n, ab, d).transpose(1, 2)).reshape(n, -1)        # [N, H*A*B]
    d_agg = torch.matmul(w.view(n, num_heads, ab).transpose(1, 2), g3).reshape(n, a_n, bd)

    row = g.row
    cnt = (g.rowptr[1:] - g.rowptr[:-1]).clamp(min=1).to(bases.dtype).view(-1, 1)
    t_sym = bases.new_zeros((n, bd))
    t_lin = bases.new_zeros((n, bd))
    t_sq = bases.new_zeros((n, bd))
    d_msg_routed = bases.new_zeros((g.nnz + 1, bd))
    vl = g.val_lin.view(-1, 1).to(bases.dtype) if g.val_lin is not None else None
    xj = bases.index_select(0, g.col)
    lin = xj if vl is None else xj * vl
    zeros = bases.new_zeros((n, bd))
    for k, a in enumerate(aggrs):
        da = d_agg[:, k]
        if a == "sum":
            t_lin = t_lin + da
        elif a == "symnorm":
            t_sym = t_sym + da
        elif a == "mean":
            t_lin = t_lin + da / cnt
        elif a in ("min", "max"):
            d_msg_routed.scatter_add_(0, args[a], da)
        elif a in ("var", "std"):
            mean = zeros.index_add(0, row, lin) / cnt
            if a == "std":
                sq = xj * xj if vl is None else xj * xj * vl
                var = zeros.index_add(0, row, sq) / cnt - mean * mean
                std = torch.sqrt(torch.relu(var) + STD_EPS)
                da = da * (var > 0).to(da.dtype) / (2 * std)
            t_sq = t_sq + da / cnt
            t_lin = t_lin - 2 * mean * da / cnt
    d_msg = d_msg_routed[: g.nnz]
    if vl is not None:
        d_msg = d_msg * vl
    lin_part = t_lin[row] + 2 * xj * t_sq[row]
    if vl is not None:
        lin_part = lin_part * vl
    d_msg = d_msg + lin_part
    if g.val_sym is not None:
        d_msg = d_msg + t_sym[row] * g.val_sym.view(-1, 1).to(bases.dtype)
    d_bases = bases.new_zeros((g.n_src, bd)).index_add_(0, g.col, d_msg)

    d_lin = d_w * w * (1 - w) if sigmoid else d_w
    return {
        "d_x": d_bases @ bases_weight.t() + d_lin @ comb_weight,
        "d_bases_weight": x.t() @ d_bases,
        "d_comb_weight": d_lin.t() @ x,
        "d_comb_bias": d_lin.sum(0),
        "d_bias": grad_out.sum(0) if bias is not None else None,
        "d_bases": d_bases, "d_weightings": d_w, "t_sym": t_sym, "t_lin": t_lin, "t_sq": t_sq,
        "agg": agg, "arg": args,
    }


# ----------------------------------------------------------------------------------------
# module mirroring the reference operator (constructor, parameters, caching, errors)
# ----------------------------------------------------------------------------------------
class EGConvOracle(torch.nn.Module):
    """Same constructor / state_dict / forward contract as the reference `EGConv` (ref :74-210)."""

    def __init__(self, in_channels, out_channels, aggrs=("symnorm",), num_heads=8, num_bases=4,
                 cached=False, add_self_loops=True, bias=True, sigmoid=False):
        super().__init__()
        if out_channels % num_heads != 0:
            raise ValueError("out_channels must be divisible by the number of heads")     # ref :89-90
        for a in aggrs:
            if a not in AGGREGATORS:
                raise ValueError("Unsupported aggregator: {}".format(a))                   # ref :92-94
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_heads, self.num_bases = num_heads, num_bases
        self.cached, self.add_self_loops = cached, add_self_loops
        self.aggregators = list(aggrs)
        self.sigmoid = sigmoid
        self.bases_weight = torch.nn.Parameter(torch.empty(in_channels, (out_channels // num_heads) * num_bases))
        self.comb_weight = torch.nn.Linear(in_channels, num_heads * num_bases * len(self.aggregators))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):                                                            # ref :117-122
        bound = (6.0 / (self.bases_weight.size(0) + self.bases_weight.size(1))) ** 0.5
        with torch.no_grad():
            self.bases_weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.zero_()
        self.comb_weight.reset_parameters()
        self._graph = None

    def prepare(self, x: Tensor, edge_index) -> OracleGraph:
        if self._graph is not None:
            return self._graph
        sym = "symnorm" in self.aggregators
        loops = self.add_self_loops
        # gcn_norm builds its unit weights with dtype=None -> torch's default dtype, whatever x is
        wdt = torch.get_default_dtype()
        if isinstance(edge_index, Tensor):
            g = graph_from_edge_index(edge_index, x.size(0), sym, loops, wdt)
        else:   # (rowptr, col, value|None) triple or anything with .csr()
            rowptr, col, value = edge_index.csr() if hasattr(edge_index, "csr") else edge_index
            g = graph_from_csr(rowptr, col, value, x.size(0), sym, loops, len(self.aggregators) > 1, wdt)
        if self.cached:
            self._graph = g
        return g

    def forward(self, x: Tensor, edge_index) -> Tensor:
        g = self.prepare(x, edge_index)
        return egconv_forward(x, g, self.bases_weight, self.comb_weight.weight, self.comb_weight.bias,
                              self.bias, self.aggregators, self.num_heads, self.sigmoid)

    def __repr__(self):
        return "{}({}, {}, {})".format("EGConv", self.in_channels, self.out_channels, self.aggregators)


# ---------------------------------------------------------------------------------------------------
# the full-graph stack  (/root/reference/experiments/mag/models.py:16-69 `EGC`)
# ---------------------------------------------------------------------------------------------------
class EGCOracle(torch.nn.Module):
    """conv -> ReLU -> dropout for all but the last layer; last layer IN->OUT_ROUNDED truncated to OUT_TRUE columns;
    log_softmax.  Same constructor arguments and `convs.{i}.*` state_dict keys as the reference class."""

    def __init__(self, hidden_channels, num_layers, dropout, num_heads, num_bases, aggrs, in_features=128,
                 out_rounded=352, out_true=349):
        super().__init__()
        dims = [in_features] + [hidden_channels] * (num_layers - 1) + [out_rounded]          # ref :22-54
        self.convs = torch.nn.ModuleList(
            EGConvOracle(dims[i], dims[i + 1], aggrs=aggrs, num_heads=num_heads, num_bases=num_bases, cached=True)
            for i in range(num_layers))
        self.dropout, self.out_true = dropout, out_true

    def forward(self, x: Tensor, adj_t) -> Tensor:
        for conv in self.convs[:-1]:                                                          # ref :61-65
            x = torch.nn.functional.dropout(torch.relu(conv(x, adj_t)), p=self.dropout, training=self.training)
        return self.convs[-1](x, adj_t)[:, :self.out_true].log_softmax(dim=-1)                # ref :68-69


# ---------------------------------------------------------------------------------------------------
# paper variant  (/root/reference/experiments/layers.py:11-228 `EfficientGraphConv` + `_AggLayer`)
# ---------------------------------------------------------------------------------------------------
PAPER_NAMES = {"symadd": "symnorm", "add": "sum", "mean": "mean", "min": "min", "max": "max", "var": "var", "std": "std"}


def paper_forward(x: Tensor, graph_in, bases_weights: Sequence[Tensor], comb_weight: Tensor, comb_bias: Tensor,
                  bias: Optional[Tensor], aggrs: Sequence[str], num_heads: int, add_self_loops: bool = True,
                  post: str = "none", graph_dtype=None) -> Tensor:
    """CPU restatement of the paper layer on top of the EGConv restatement above.
    `graph_in`: edge_index [2, E] or (rowptr, col, value) of adj_t.  `post`: none | softmax | sigmoid | hardtanh.
    Self-loops / gcn_norm only for `symadd` (layers.py:167-188); every other aggregator sees the raw graph;
    comb-weight columns are ordered h * (B * A) + b * A + a (layers.py:106-129)."""
    n, b, a = x.size(0), len(bases_weights), len(aggrs)
    gdt = graph_dtype or x.dtype          # gcn_norm on a value-less SparseTensor materialises fp32 ones in the reference
    bases = torch.cat([x @ w for w in bases_weights], dim=1)                     # layers.py:97-101
    lin = x @ comb_weight.t() + comb_bias                                        # layers.py:109
    if post == "softmax":
        w = lin.view(n, num_heads, b * a).softmax(dim=-1)                        # layers.py:112-120
    elif post == "sigmoid":
        w = torch.sigmoid(lin)
    elif post == "hardtanh":
        w = torch.nn.functional.hardtanh(lin)
    else:
        w = lin
    w = w.reshape(n, num_heads, b, a)

    def make_graph(symnorm: bool) -> OracleGraph:
        loops = add_self_loops and symnorm
        if isinstance(graph_in, Tensor):
            return graph_from_edge_index(graph_in, n, symnorm, loops, dtype=gdt)
        rowptr, col, value = graph_in
        return graph_from_csr(rowptr, col, value, n, symnorm, loops, False, dtype=gdt)

    cols = []
    for name in aggrs:                                                           # one _AggLayer per aggregator
        agg, _ = aggregate(make_graph(name == "symadd"), bases, [PAPER_NAMES[name]])
        cols.append(agg[:, 0].reshape(n, b, -1))                                 # [N, B, D]
    y = torch.stack(cols, dim=2)                                                 # [N, B, A, D]  layers.py:108
    z = (w.unsqueeze(-1) * y.unsqueeze(1)).sum(dim=(2, 3)).reshape(n, -1)        # layers.py:131-135
    return z + bias if bias is not None else z


def arxivnet_forward(x: Tensor, edge_index: Tensor, state: dict, num_layers: int, heads: int, bases: int,
                     aggrs: Sequence[str], residual: bool, training: bool = False, eps: float = 1e-5,
                     graph_dtype=torch.float32) -> Tensor:
    """CPU restatement of the reference's normalised full-graph model (dropout 0):
    /root/reference/experiments/arxiv/norm_models.py:31-43 - embed (one Linear: experiments/utils.py:33-43 with two sizes)
    -> per layer [paper conv -> BatchNorm1d -> ReLU -> (+ identity)] -> Linear -> log_softmax.  `state` = the reference's
    state_dict (any float dtype).  training=True uses the batch statistics (biased variance, as BatchNorm1d normalises).
    graph_dtype: gcn_norm of an edge_index input materialises its unit edge weights in the default dtype (fp32) whatever
    the model's dtype, so the reference's symnorm weights are fp32 values even in an fp64 run."""
    h = x @ state["embed.0.weight"].t() + state["embed.0.bias"]                                 # norm_models.py:32
    for i in range(num_layers):
        identity = h
        w_b = [state[f"convs.{i}.bases_weight.{b}"] for b in range(bases)]
        y = paper_forward(h, edge_index, w_b, state[f"convs.{i}.comb_weights.weight"], state[f"convs.{i}.comb_weights.bias"],
                          state.get(f"convs.{i}.bias"), aggrs, heads, graph_dtype=graph_dtype)  # :35
        if training:
            mean, var = y.mean(0), y.var(0, unbiased=False)
        else:
            mean, var = state[f"bns.{i}.running_mean"], state[f"bns.{i}.running_var"]
        y = (y - mean) / torch.sqrt(var + eps) * state[f"bns.{i}.weight"] + state[f"bns.{i}.bias"]   # :36
        y = torch.relu(y)                                                                      # :37 (dropout 0: identity)
        h = y + identity if residual else y                                                    # :39-40
    return (h @ state["out.weight"].t() + state["out.bias"]).log_softmax(dim=-1)               # :42-43
