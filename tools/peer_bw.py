"""NVLink push microbenchmark (run under torchrun, 2+ ranks):  egc_peer_push_rows of `mb` MB of rows into every peer,
one direction (only rank 0 pushes) and all directions at once.  EGC_PEER_PUSH_CTAS_PER_SM selects the grid.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_bw.py [width] [mb]"""
import ctypes, os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from egc_b200 import _lib
from egc_b200.peer import PeerSegment, _ptr_array

width = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
rows_per_peer = mb * (1 << 20) // (width * 4) // max(world - 1, 1)
n_rows = rows_per_peer * (world - 1)
seg = PeerSegment([("recv", 4 * n_rows * width)], dev)
src = torch.randn(2 * n_rows, width, device=dev)
peers = [q for q in range(world) if q != rank]
seg_ptr = (ctypes.c_int32 * (len(peers) + 1))()
for s in range(len(peers)):
    seg_ptr[s + 1] = seg_ptr[s] + rows_per_peer
# every peer q receives my rows at its slot for sender = me (slots ordered by sender rank, skipping the receiver)
dst = _ptr_array([seg.peer_ptr(q, "recv", (rank if rank < q else rank - 1) * rows_per_peer * width * 4) for q in peers])
srcs = _ptr_array([src.data_ptr()] * len(peers))
idx_seq = torch.arange(n_rows, device=dev, dtype=torch.int32)
idx_rand = torch.randperm(2 * n_rows, device=dev)[:n_rows].to(torch.int32).sort().values
st = lambda: torch.cuda.current_stream().cuda_stream


def push(index):
    _lib.check(lib.egc_peer_push_rows(len(peers), srcs, dst, seg_ptr, None if index is None else index.data_ptr(), width, None,
                                      world, rank, 0, None, None, st()), "push")


def timed(fn, active, reps=10):
    for _ in range(3):
        if active:
            fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if active:
            fn()
    e1.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


gb = n_rows * width * 4 / 1e9
for name, index in (("contiguous rows", None), ("sorted random rows (gather)", idx_rand)):
    t1 = timed(lambda: push(index), rank == 0)
    ta = timed(lambda: push(index), True)
    if rank == 0:
        print(f"width {width} floats, {gb * 1e3:.0f} MB per rank, {world} ranks, ctas/SM {os.environ.get('EGC_PEER_PUSH_CTAS_PER_SM', 'default')}: {name}: "
              f"one sender {t1 * 1e3:.0f} us = {gb / t1 * 1e3:.0f} GB/s; all senders {ta * 1e3:.0f} us = {gb / ta * 1e3:.0f} GB/s per direction", flush=True)
# reference: cudaMemcpy peer (torch copy_ between mapped buffers is not available; use NCCL send/recv as a yardstick)
buf = torch.empty(n_rows * width, device=dev)
if world == 2:
    def sr():
        if rank == 0:
            dist.send(src.view(-1)[: buf.numel()], 1)
        else:
            dist.recv(buf, 0)
    t = timed(sr, True)
    if rank == 0:
        print(f"NCCL send/recv of the same bytes: {t * 1e3:.0f} us = {gb / t * 1e3:.0f} GB/s", flush=True)
seg.close()
dist.destroy_process_group()
