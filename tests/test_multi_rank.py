"""N>1 host logic on CPU: gloo, world_size 2 -- table blob broadcast from rank 0 and disjoint frame sharding."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from mercury_b200.dist import shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 65536, 1048576 + 3):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist

    import mercury_b200 as mb
    from mercury_b200.dist import broadcast_tables, shard_range as sr
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = mb.build_tables_host() if rank == 0 else None
    got = broadcast_tables(blob, src=0)
    a, b = sr(1001, rank, world)
    x, pl = mb.synth_frames(8, 1001, seed=3, esn0_db=300.0, n_threads=1) if rank == 0 else (None, None)
    q.put((rank, int(got.size), int(np.frombuffer(got.tobytes(), np.uint32)[0]), __import__("zlib").crc32(got.tobytes()), a, b))
    dist.barrier()
    dist.destroy_process_group()


def test_table_broadcast_gloo_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, n0, magic0, h0, a0, b0), (r1, n1, magic1, h1, a1, b1) = res
    assert n0 == n1 > 100000 and magic0 == magic1 == 0x42324D42 and h0 == h1
    assert (a0, b0, a1, b1) == (0, 501, 501, 1001)
