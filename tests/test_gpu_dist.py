"""Multi-GPU parity: the row-partitioned layer over 2 ranks (NCCL) == the single-GPU layer.
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


def _worker(rank, world, port, n, ei, x, go, ret):
    import egc_b200
    from egc_b200.dist import PartitionedGraph, partitioned_egconv
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.manual_seed(0)
        conv = egc_b200.EGConv(64, 128, aggrs=AGGRS, num_heads=4, num_bases=4).to(dev)
        params = list(conv.parameters())
        # single-GPU reference on this rank's device
        xs = x.to(dev).requires_grad_(True)
        out_ref = conv(xs, ei.to(dev))
        grads_ref = torch.autograd.grad(out_ref, [xs] + params, go.to(dev))
        g = egc_b200.GraphStructure.from_edge_index(ei.to(dev), n, True, True)
        pg = PartitionedGraph.from_global(g, rank, world, dev)
        b, e = pg.part.row_begin, pg.part.row_end
        xl = x[b:e].to(dev).requires_grad_(True)
        out = partitioned_egconv(xl, pg, conv)
        grads = torch.autograd.grad(out, [xl] + params, go[b:e].to(dev))
        assert rel_err(out, out_ref[b:e]) < 1e-6
        assert rel_err(grads[0], grads_ref[0][b:e]) < 1e-5
        for a, r in zip(grads[1:], grads_ref[1:]):
            assert rel_err(a, r) < 1e-5
        assert pg.part.interior_rows.numel() > 0 and pg.part.n_halo > 0
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_partitioned_layer_matches_single_gpu_nccl_world2():
    n = 4000
    ei = random_graph(n, 30000, seed=5, hub=900)
    blk = torch.randint(0, n // 2, (2, 15000))
    ei = torch.cat([ei, blk, blk + n // 2], 1)
    torch.manual_seed(1)
    x, go = torch.randn(n, 64), torch.randn(n, 128)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, ei, x, go, ret), nprocs=2, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}
