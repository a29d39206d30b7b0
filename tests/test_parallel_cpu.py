"""CPU, world_size 2, gloo: the N>1 host protocol (batch sharding, blob broadcast, result gather)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import csfm_b200
    par = csfm_b200.parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1) index replication protocol: only rank 0 knows the blob and its size
        rng = np.random.default_rng(1)
        blob = torch.from_numpy(rng.integers(0, 256, 100_003, dtype=np.uint8))
        got = par.broadcast_bytes(blob if rank == 0 else None, src=0)
        assert got.dtype == torch.uint8 and torch.equal(got, blob)
        # 2) batch sharding: ragged patterns, every query lands on exactly one rank, order preserved
        pats = [bytes(rng.integers(65, 70, int(m), dtype=np.uint8)) for m in rng.integers(0, 9, 1001)]
        data, offs = csfm_b200.pack_patterns(pats)
        d, o, lo, hi = par.shard_patterns(data, offs, rank, world)
        assert o[0] == 0 and len(o) == hi - lo + 1
        mine = [bytes(d[int(o[k]):int(o[k + 1])]) for k in range(hi - lo)]
        assert mine == pats[lo:hi]
        # 3) gather: a stand-in "count" (pattern length) computed on the shard, gathered everywhere
        local = torch.tensor([len(p) for p in mine], dtype=torch.int64)
        allc = par.gather_counts(local, len(pats))
        assert allc.tolist() == [len(p) for p in pats]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    import csfm_b200
    for total in (0, 1, 7, 8, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [csfm_b200.parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(120)
def test_two_rank_protocol_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
