"""Row-partition + halo-exchange logic on CPU with world_size = 2 (gloo).  The compute on each rank is the
CPU oracle (test infrastructure), so what is checked is exactly the product's index / communication
plumbing (`egc_b200.dist.PartitionPlan`, `HaloExchange`): partitioned result == single-process result."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from egc_b200.dist import HaloExchange, PartitionPlan, balanced_row_bounds, n_target_streams, transpose_csr
from oracle import restatement as R
from tests.util import random_graph, rel_err

AGGRS = ["symnorm", "max", "std", "mean"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _single_process(n, ei, x, go, seed):
    torch.manual_seed(seed)
    layer = R.EGConvOracle(12, 32, aggrs=AGGRS, num_heads=4, num_bases=4).double()
    g = layer.prepare(x, ei)
    xx = x.clone().requires_grad_(True)
    out = layer(xx, ei)
    grads = torch.autograd.grad(out, [xx] + list(layer.parameters()), go)
    return layer, g, out.detach(), grads


def _worker(rank, world, port, n, ei, x, go, seed, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        layer, g, out_ref, grads_ref = _single_process(n, ei, x, go, seed)
        plan = PartitionPlan(g.rowptr, g.col, world, val_sym=g.val_sym)
        part = plan.local(rank)
        ex = HaloExchange(part, "cpu")
        b, e = part.row_begin, part.row_end
        x_loc = x[b:e].clone().requires_grad_(True)
        # forward: local projection, halo exchange of basis rows, aggregation over [own | halo]
        bases_loc, w_loc = R.project(x_loc, layer.bases_weight, layer.comb_weight.weight, layer.comb_weight.bias)
        halo = ex.forward(bases_loc.detach())
        # (a row-sliced matmul may block differently from the full one: equal up to fp64 rounding, not bit-equal)
        assert rel_err(halo, (x @ layer.bases_weight.detach())[part.halo_ids]) < 1e-13
        halo = halo.requires_grad_(True)
        local_graph = R.OracleGraph(part.rowptr, part.col, part.n_local, part.n_local + part.n_halo, val_sym=part.val_sym)
        agg, _ = R.aggregate(local_graph, torch.cat([bases_loc, halo]), AGGRS)
        out_loc = R.combine(w_loc, agg, layer.bias, 4)
        assert rel_err(out_loc, out_ref[b:e]) < 1e-12
        # backward: local autograd gives partial sums for halo sources; they travel back to the owners
        params = list(layer.parameters())
        grads = torch.autograd.grad(out_loc, [x_loc, halo, bases_loc] + params, go[b:e], allow_unused=True,
                                    retain_graph=True)
        d_halo = grads[1]
        extra = torch.zeros_like(bases_loc)
        ex.reverse(d_halo, extra)                                        # contributions from the other rank
        d_x_extra, *d_par_extra = torch.autograd.grad(bases_loc, [x_loc] + params, extra, allow_unused=True)
        d_x = grads[0] + d_x_extra
        assert rel_err(d_x, grads_ref[0][b:e]) < 1e-11
        for gp, ge, gr in zip(grads[3:], d_par_extra, grads_ref[1:]):
            tot = gp + (ge if ge is not None else 0)
            dist.all_reduce(tot)
            assert rel_err(tot, gr) < 1e-11
        # interior rows only touch local sources
        rows = torch.repeat_interleave(torch.arange(part.n_local), part.rowptr[1:] - part.rowptr[:-1])
        touched = torch.zeros(part.n_local, dtype=torch.bool)
        touched[rows[part.col >= part.n_local]] = True
        assert torch.equal(torch.nonzero(~touched).flatten(), part.interior_rows)
        assert part.interior_rows.numel() + part.boundary_rows.numel() == part.n_local
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_partitioned_layer_matches_single_process_gloo_world2():
    torch.manual_seed(0)
    n = 240
    ei = random_graph(n, 1500, seed=5, hub=90)
    # add locality so that interior rows exist
    blk = torch.randint(0, n // 2, (2, 600))
    ei = torch.cat([ei, blk, blk + n // 2], 1)
    x = torch.randn(n, 12, dtype=torch.float64)
    go = torch.randn(n, 32, dtype=torch.float64)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n, ei, x, go, 3, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_partition_plan_invariants(world):
    n = 500
    ei = random_graph(n, 4000, seed=9, hub=300)
    g = R.graph_from_edge_index(ei, n, True, True)
    plan = PartitionPlan(g.rowptr, g.col, world, val_sym=g.val_sym)
    assert plan.bounds[0] == 0 and plan.bounds[-1] == n and sorted(plan.bounds) == plan.bounds
    nnz = [int(g.rowptr[plan.bounds[r + 1]] - g.rowptr[plan.bounds[r]]) for r in range(world)]
    assert max(nnz) <= g.nnz / world + int((g.rowptr[1:] - g.rowptr[:-1]).max()) + n     # balanced up to one row
    total_rows = 0
    for r in range(world):
        part = plan.local(r)
        total_rows += part.n_local
        ext = torch.cat([torch.arange(part.row_begin, part.row_end), part.halo_ids])
        lo, hi = int(g.rowptr[part.row_begin]), int(g.rowptr[part.row_end])
        assert torch.equal(ext[part.col], g.col[lo:hi])                 # remapped columns point at the same nodes
        assert torch.equal(part.val_sym, g.val_sym[lo:hi])
        assert sum(part.recv_counts) == part.n_halo and part.recv_counts[r] == 0
        for q in range(world):                                          # what q sends me is what I asked of q
            assert torch.equal(plan.local(q).send_rows[r] + plan.bounds[q], plan._needs[r][q])
    assert total_rows == n
    assert balanced_row_bounds(g.rowptr, 1) == [0, n]


# ------------------------------------------------------------------------------------------------
# T exchange: the backward of layers without min / max exchanges the target-side stream rows over the TRANSPOSED plan
# ------------------------------------------------------------------------------------------------
def _directed_graph(n, seed):
    """NOT symmetric: the transposed plan's halo differs from the forward one."""
    ei = random_graph(n, 1400, seed=seed, hub=80)
    gen = torch.Generator().manual_seed(seed + 1)
    extra = torch.stack([torch.randint(0, n // 3, (500,), generator=gen), torch.randint(n // 2, n, (500,), generator=gen)])
    return torch.cat([ei, extra], 1)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_transposed_plan_invariants(world):
    n = 400
    g = R.graph_from_edge_index(_directed_graph(n, 21), n, True, True)
    plan = PartitionPlan(g.rowptr, g.col, world, val_sym=g.val_sym)
    plan_t = plan.transposed()
    assert plan_t.bounds == plan.bounds
    rowptr_t, col_t, val_t = transpose_csr(g.rowptr, g.col, g.val_sym)
    dense = torch.zeros(n, n, dtype=torch.float64)
    rows = torch.repeat_interleave(torch.arange(n), g.rowptr[1:] - g.rowptr[:-1])
    dense.index_put_((rows, g.col), g.val_sym.double(), accumulate=True)
    rows_t = torch.repeat_interleave(torch.arange(n), rowptr_t[1:] - rowptr_t[:-1])
    dense_t = torch.zeros(n, n, dtype=torch.float64)
    dense_t.index_put_((rows_t, col_t), val_t.double(), accumulate=True)
    assert torch.equal(dense_t, dense.t())
    for r in range(world):
        tp, p = plan_t.local(r), plan.local(r)
        assert (tp.row_begin, tp.row_end) == (p.row_begin, p.row_end)
        ext = torch.cat([torch.arange(tp.row_begin, tp.row_end), tp.halo_ids])
        lo, hi = int(rowptr_t[tp.row_begin]), int(rowptr_t[tp.row_end])
        assert torch.equal(ext[tp.col], col_t[lo:hi]) and torch.equal(tp.val_sym, val_t[lo:hi])
        # my transposed halo = the remote TARGETS of my source columns = the ranks whose forward halo contains my rows
        for q in range(world):
            assert torch.equal(plan_t.local(q).send_rows[r] + plan.bounds[q], plan_t._needs[r][q])
    for r in range(world):                                      # own-column / halo-column parts of a local block
        for part in (plan.local(r), plan_t.local(r)):
            (rp_a, col_a, sym_a, _), (rp_b, col_b, sym_b, _) = part.split_by_column()
            assert bool((col_a < part.n_local).all()) and bool((col_b >= part.n_local).all())
            assert torch.equal(rp_a + rp_b, part.rowptr) and col_a.numel() + col_b.numel() == part.col.numel()
            dense = lambda rp, c, v: torch.zeros(part.n_local, part.n_local + part.n_halo, dtype=torch.float64).index_put_(  # noqa: E731
                (torch.repeat_interleave(torch.arange(part.n_local), rp[1:] - rp[:-1]), c), v.double(), accumulate=True)
            assert torch.equal(dense(rp_a, col_a, sym_a) + dense(rp_b, col_b, sym_b), dense(part.rowptr, part.col, part.val_sym))
    assert n_target_streams(["symnorm"]) == 1 and n_target_streams(["sum", "mean"]) == 1
    assert n_target_streams(["symnorm", "mean"]) == 2 and n_target_streams(["symnorm", "max", "std"]) == 3


def _t_worker(rank, world, port, n, ei, x, go, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        aggrs = ["symnorm"]
        torch.manual_seed(7)
        layer = R.EGConvOracle(12, 32, aggrs=aggrs, num_heads=4, num_bases=4).double()
        g = layer.prepare(x, ei)
        # single-process truth: gradient of the loss with respect to the projected bases
        bases_full, w_full = R.project(x, layer.bases_weight, layer.comb_weight.weight, layer.comb_weight.bias)
        bases_full = bases_full.detach().requires_grad_(True)
        agg_full, _ = R.aggregate(g, bases_full, aggrs)
        out_full = R.combine(w_full, agg_full, layer.bias, 4)
        (d_bases_ref,) = torch.autograd.grad(out_full, [bases_full], go)
        plan = PartitionPlan(g.rowptr, g.col, world, val_sym=g.val_sym)
        part, tpart = plan.local(rank), plan.transposed().local(rank)
        b, e = part.row_begin, part.row_end
        # pass 1 on this rank: the target-side stream of my rows (symnorm: t_sym = d(loss) / d(agg), [n_local, B*D])
        halo = HaloExchange(part, "cpu").forward(bases_full.detach()[b:e])
        local_graph = R.OracleGraph(part.rowptr, part.col, part.n_local, part.n_local + part.n_halo, val_sym=part.val_sym)
        bases_ext = torch.cat([bases_full.detach()[b:e], halo])
        agg = R.aggregate(local_graph, bases_ext, aggrs)[0].detach().requires_grad_(True)
        out_loc = R.combine(w_full[b:e], agg, layer.bias, 4)
        (t_own,) = torch.autograd.grad(out_loc, [agg], go[b:e])
        t_own = t_own.reshape(part.n_local, -1)
        # T exchange over the transposed plan, then the column pass over my own columns
        t_ext = torch.cat([t_own, HaloExchange(tpart, "cpu").forward(t_own)])
        rows = torch.repeat_interleave(torch.arange(tpart.n_local), tpart.rowptr[1:] - tpart.rowptr[:-1])
        d_bases = torch.zeros(tpart.n_local, t_own.size(1), dtype=torch.float64)
        d_bases.index_add_(0, rows, tpart.val_sym.double().unsqueeze(1) * t_ext[tpart.col])
        assert rel_err(d_bases, d_bases_ref[b:e]) < 1e-12
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_t_exchange_backward_matches_single_process_gloo_world2():
    n = 300
    ei = _directed_graph(n, 31)
    torch.manual_seed(1)
    x = torch.randn(n, 12, dtype=torch.float64)
    go = torch.randn(n, 32, dtype=torch.float64)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_t_worker, args=(2, _free_port(), n, ei, x, go, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


# ------------------------------------------------------------------------------------------------
# data-parallel mini-batches (BASELINE configs 1 / 5): weighted flat all-reduce of the parameter gradients
# ------------------------------------------------------------------------------------------------
def _dp_model(seed):
    torch.manual_seed(seed)
    return torch.nn.ModuleList([R.EGConvOracle(6, 16, aggrs=["sum"], num_heads=4, num_bases=4),
                                R.EGConvOracle(16, 16, aggrs=["symnorm", "max", "std"], num_heads=4, num_bases=4)]).double()


def _dp_loss(model, graphs):
    from oracle import batching as OB
    x, ei, batch, ptr = OB.collate(graphs)
    h = x.double()
    for layer in model:
        h = torch.relu(layer(h, ei))
    return OB.global_pool(h, batch, len(graphs), "mean").pow(2).sum(1).mean()      # mean over the graphs


def _dp_graphs():
    from oracle import batching as OB
    gen = torch.Generator().manual_seed(4)
    return [(torch.randn(n, 6, generator=gen), ei, n) for _, ei, n in OB.zinc_like_graphs(10, seed=2)]


def _dp_worker(rank, world, port, ret):
    from egc_b200.dist import GradientAllReduce
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        graphs = _dp_graphs()
        ref = _dp_model(7)
        _dp_loss(ref, graphs).backward()
        mine = graphs[:3] if rank == 0 else graphs[3:]                 # unequal shares: 3 and 7 graphs
        model = _dp_model(7)
        _dp_loss(model, mine).backward()
        sync = GradientAllReduce(model.parameters(), bucket_bytes=2048)   # several buckets
        assert len(sync.buckets) > 1
        sync(weight=len(mine) / len(graphs))
        for p, q in zip(model.parameters(), ref.parameters()):
            assert rel_err(p.grad, q.grad) < 1e-11
        # the same through the overlapped variant: per-layer buckets launched from inside backward, two steps in a row
        from egc_b200.dist import OverlappedGradientAllReduce
        model2 = _dp_model(7)
        over = OverlappedGradientAllReduce(list(model2))
        assert len(over.buckets) == 2
        for _ in range(2):
            for p in model2.parameters():
                p.grad = None
            over.begin(weight=len(mine) / len(graphs))
            _dp_loss(model2, mine).backward()
            over.finish()
            for p, q in zip(model2.parameters(), ref.parameters()):
                assert rel_err(p.grad, q.grad) < 1e-11
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradients_match_single_process_gloo_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_every_rank_owns_a_row_even_with_a_dominant_hub():
    """A hub row heavier than a whole share must not leave a rank without rows (empty ranks block the step)."""
    from egc_b200.dist import PartitionPlan, balanced_row_bounds
    n = 12
    counts = torch.ones(n, dtype=torch.long)
    counts[0] = 10_000                                        # one hub row outweighs everything else
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
    for world in (2, 4, 8):
        b = balanced_row_bounds(rowptr, world)
        assert b[0] == 0 and b[-1] == n and all(b[i + 1] > b[i] for i in range(world)), b
    with pytest.raises(ValueError):
        PartitionPlan(torch.tensor([0, 1, 2]), torch.tensor([0, 1]), 4)
