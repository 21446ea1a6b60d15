"""Multi-GPU parity: the row-partitioned layer over 2 ranks == the single-GPU layer, for both transports
(our NVLink peer-memory kernels and the NCCL baseline), across repeated steps, under no_grad, for a two-layer
stack, and when the whole step is replayed from a CUDA graph.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import random_graph, rel_err

pytestmark = pytest.mark.gpu
AGGRS = ["symnorm", "max", "std"]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def check_two_layers(conv, conv2, params2, pg, x, go, eid, b, e, dev):
    from egc_b200.dist import partitioned_egconv
    xs = x.to(dev).requires_grad_(True)
    ref2 = conv2(torch.relu(conv(xs, eid)), eid)
    gref2 = torch.autograd.grad(ref2, [xs] + params2, go.to(dev))
    xl = x[b:e].to(dev).requires_grad_(True)
    out2 = partitioned_egconv(torch.relu(partitioned_egconv(xl, pg, conv)), pg, conv2)
    g2 = torch.autograd.grad(out2, [xl] + params2, go[b:e].to(dev))
    assert rel_err(out2, ref2[b:e]) < 1e-5
    assert rel_err(g2[0], gref2[0][b:e]) < 2e-5
    for a, r in zip(g2[1:], gref2[1:]):
        assert rel_err(a, r) < 2e-5


def _worker(rank, world, port, n, ei, x, go, transport, ret, aggrs=None, expect_t=False, env=None):
    AGGRS = aggrs or globals()["AGGRS"]
    os.environ.update(env or {})
    import egc_b200
    from egc_b200.dist import GraphedStep, PartitionedGraph, partitioned_egconv
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    pg = None
    try:
        torch.manual_seed(0)
        conv = egc_b200.EGConv(64, 128, aggrs=AGGRS, num_heads=4, num_bases=4).to(dev)
        conv2 = egc_b200.EGConv(128, 128, aggrs=AGGRS, num_heads=4, num_bases=4).to(dev)
        params = list(conv.parameters())
        params2 = params + list(conv2.parameters())
        eid = ei.to(dev)
        g = egc_b200.GraphStructure.from_edge_index(eid, n, True, True)
        pg = PartitionedGraph.from_global(g, rank, world, dev, transport=transport)
        assert pg.transport == transport
        assert pg.uses_t_exchange(AGGRS) == expect_t
        if env and "EGC_DIST_OVERLAP" in env:
            want = (env["EGC_DIST_OVERLAP"] == "1" and transport == "peer" and expect_t and
                    all(a in ("sum", "symnorm") for a in AGGRS))
            assert pg.uses_overlap(AGGRS) == want
        b, e = pg.part.row_begin, pg.part.row_end
        assert pg.part.interior_rows.numel() > 0 and pg.part.n_halo > 0

        def check_step(x_full, go_full, tag):
            xs = x_full.to(dev).requires_grad_(True)
            out_ref = conv(xs, eid)
            grads_ref = torch.autograd.grad(out_ref, [xs] + params, go_full.to(dev))
            xl = x_full[b:e].to(dev).requires_grad_(True)
            out = partitioned_egconv(xl, pg, conv)
            grads = torch.autograd.grad(out, [xl] + params, go_full[b:e].to(dev))
            assert rel_err(out, out_ref[b:e]) < 1e-6, tag
            assert rel_err(grads[0], grads_ref[0][b:e]) < 1e-5, tag
            for a, r in zip(grads[1:], grads_ref[1:]):
                assert rel_err(a, r) < 1e-5, tag

        # repeated steps with fresh data: exercises the epoch flags and the buffer-reuse ordering
        for it in range(3):
            gen = torch.Generator().manual_seed(10 + it)
            check_step(torch.randn(x.shape, generator=gen), torch.randn(go.shape, generator=gen), f"step {it}")
        pg.check()
        # inference under no_grad, twice in a row, then a training step again
        with torch.no_grad():
            for _ in range(2):
                out_ref = conv(x.to(dev), eid)
                out = partitioned_egconv(x[b:e].to(dev), pg, conv)
                assert rel_err(out, out_ref[b:e]) < 1e-6
        pg.check()
        check_step(x, go, "after no_grad")
        pg.check()

        def two_layers():      # two stacked layers (each keeps its own exchange buffers)
            check_two_layers(conv, conv2, params2, pg, x, go, eid, b, e, dev)

        two_layers()
        pg.check()
        import gc
        gc.collect()           # no autograd graph of the eager steps may stay alive across the capture below

        if transport == "peer":
            # the whole step (exchange kernels and flags included) replayed from a CUDA graph
            x_static = x[b:e].to(dev).requires_grad_(True)
            go_static = go[b:e].to(dev)

            def step():
                o = partitioned_egconv(x_static, pg, conv)
                return (o,) + torch.autograd.grad(o, [x_static] + params, go_static)

            graphed = GraphedStep(step, warmup=2)
            pg.check()
            for it in range(3):
                gen = torch.Generator().manual_seed(50 + it)
                xf, gf = torch.randn(x.shape, generator=gen), torch.randn(go.shape, generator=gen)
                with torch.no_grad():
                    x_static.copy_(xf[b:e].to(dev))
                    go_static.copy_(gf[b:e].to(dev))
                res = graphed.replay()
                xs = xf.to(dev).requires_grad_(True)
                out_ref = conv(xs, eid)
                grads_ref = torch.autograd.grad(out_ref, [xs] + params, gf.to(dev))
                assert rel_err(res[0], out_ref[b:e]) < 1e-6
                assert rel_err(res[1], grads_ref[0][b:e]) < 1e-5
                for a, r in zip(res[2:], grads_ref[1:]):
                    assert rel_err(a, r) < 1e-5
            pg.check()
        torch.cuda.synchronize()
        ret[rank] = "ok"
    finally:
        if pg is not None:
            try:
                pg.close()
            except Exception:
                pass
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_partitioned_layer_matches_single_gpu_world2(transport):
    n = 4000
    ei = random_graph(n, 30000, seed=5, hub=900)
    blk = torch.randint(0, n // 2, (2, 15000))
    ei = torch.cat([ei, blk, blk + n // 2], 1)
    torch.manual_seed(1)
    x, go = torch.randn(n, 64), torch.randn(n, 128)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, ei, x, go, transport, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("transport", ["peer", "nccl"])
@pytest.mark.parametrize("aggrs,overlap", [(["symnorm"], "1"), (["symnorm"], "0"), (["sum"], "1"), (["mean"], "1"),
                                           (["sum", "symnorm"], "1")], ids=lambda a: "+".join(a) if isinstance(a, list) else f"overlap{a}")
def test_partitioned_layer_with_t_exchange_matches_single_gpu_world2(transport, aggrs, overlap):
    """Layers without min / max and one target-side stream (EGC-S) exchange the stream rows over the TRANSPOSED plan and
    run the column pass over their own columns only.  The graph is NOT symmetric (extra one-way edges from the low to the
    high node range), so the transposed halo differs from the forward one.  overlap = 1 (peer transport, sum / symnorm): the
    exchanges run on the copy engine next to the launch over the own-column entries, a second launch continues the sums."""
    n = 4000
    ei = random_graph(n, 30000, seed=6, hub=900)
    blk = torch.randint(0, n // 2, (2, 15000))
    one_way = torch.stack([torch.randint(0, n // 3, (4000,)), torch.randint(n // 2, n, (4000,))])
    ei = torch.cat([ei, blk, blk + n // 2, one_way], 1)
    torch.manual_seed(2)
    x, go = torch.randn(n, 64), torch.randn(n, 128)
    mgr = mp.Manager()
    ret = mgr.dict()
    expect_t = len(aggrs) == 1                                  # sum + symnorm: two streams, partial-sum exchange
    mp.spawn(_worker, args=(2, _free_port(), n, ei, x, go, transport, ret, aggrs, expect_t, {"EGC_DIST_OVERLAP": overlap}),
             nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


# ------------------------------------------------------------------------------------------------
# data-parallel mini-batches (BASELINE config 5: CIFAR-shaped EGC-M batches, gradient all-reduce)
# ------------------------------------------------------------------------------------------------
def _dp_model(dev):
    import egc_b200
    torch.manual_seed(11)
    return torch.nn.ModuleList([egc_b200.EGConv(5, 128, aggrs=AGGRS, num_heads=4, num_bases=4),
                                egc_b200.EGConv(128, 128, aggrs=AGGRS, num_heads=4, num_bases=4)]).to(dev)


def _dp_loss(model, graphs, dev):
    import egc_b200
    b = egc_b200.collate(graphs, device=dev)
    h = b.x
    for layer in model:
        h = torch.relu(layer(h, b.edge_index))
    return egc_b200.global_mean_pool(h, b).pow(2).sum(1).mean()


def _dp_worker(rank, world, port, ret):
    from egc_b200.dist import GradientAllReduce
    from oracle import batching as OB
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        graphs = OB.cifar_like_graphs(24, seed=3)
        ref = _dp_model(dev)
        _dp_loss(ref, graphs, dev).backward()
        mine = graphs[:9] if rank == 0 else graphs[9:]
        model = _dp_model(dev)
        _dp_loss(model, mine, dev).backward()
        GradientAllReduce(model.parameters())(weight=len(mine) / len(graphs))
        torch.cuda.synchronize()
        for p, q in zip(model.parameters(), ref.parameters()):
            assert rel_err(p.grad, q.grad) < 2e-5
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_data_parallel_gradients_match_single_gpu_world2():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_dp_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
