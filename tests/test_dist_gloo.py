"""CPU, world_size 2 over gloo: the N>1 host logic (sharding, index broadcast, max-over-ranks timing)."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from nvbio_b200.dist import shard_range
    for n in (0, 1, 7, 8, 9, 1000003):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nvbio_b200 import dist as nd
    from nvbio_b200.fmindex import FMIndexDevice

    class FakeIndex:                       # FMIndexDevice asserts CUDA tensors; mimic its fields on CPU
        pass
    dev = torch.device("cpu")
    if rank == 0:
        f = FakeIndex()
        f.length, f.primary, f.L2, f.sa_interval = 1000, 17, [0, 250, 500, 760, 1000], 4
        f.bwt_occ = torch.arange(16 * 8, dtype=torch.int32)
        f.ssa = torch.arange(63, dtype=torch.int32)
        genome = torch.arange(70, dtype=torch.int32)
    else:
        f, genome = None, None
    # patch the constructor check so that the CPU tensors are accepted in this gloo test
    orig = FMIndexDevice.__init__

    def init(self, bwt_occ, ssa, L2, length, primary, sa_interval=16):
        self.bwt_occ, self.ssa, self.L2, self.length, self.primary = bwt_occ, ssa, [int(v) for v in L2], int(length), int(primary)
        self.sa_interval = sa_interval
    FMIndexDevice.__init__ = init
    g, gen = nd.broadcast_index(f, genome, dev, src=0)
    FMIndexDevice.__init__ = orig
    ok = (g.length == 1000 and g.primary == 17 and g.sa_interval == 4 and list(g.L2) == [0, 250, 500, 760, 1000] and
          torch.equal(g.bwt_occ, torch.arange(16 * 8, dtype=torch.int32)) and torch.equal(g.ssa, torch.arange(63, dtype=torch.int32)) and
          torch.equal(gen, torch.arange(70, dtype=torch.int32)))
    t = nd.max_over_ranks(1.0 + rank, dev)
    s = nd.sum_over_ranks(10.0 * (rank + 1), dev)
    b, e = nd.shard_range(101, rank, world)
    q.put((rank, ok, t, s, b, e))
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_index_and_reductions_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    assert all(abs(r[2] - 2.0) < 1e-9 and abs(r[3] - 30.0) < 1e-9 for r in res)
    assert (res[0][4], res[0][5], res[1][4], res[1][5]) == (0, 51, 51, 101)
