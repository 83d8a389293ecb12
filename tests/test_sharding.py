"""CPU: the N>1 path -- contiguous work-balanced shards, host-side ordered gather -- exercised with
world_size 2 and 3 over the gloo backend (no GPU; each rank decodes its shard with the oracle)."""
import os
import socket

import numpy as np
import pytest

import dnab_testutil as util
from dnastore_b200 import sharding


def test_shard_bounds_cover_and_balance():
    rng = np.random.default_rng(1)
    for n, world in [(0, 2), (1, 4), (7, 2), (1000, 8), (1000, 3)]:
        lens = rng.integers(0, 300, size=n)
        cuts = sharding.shard_bounds(lens, world)
        assert cuts[0] == 0 and cuts[-1] == n and len(cuts) == world + 1
        assert (np.diff(cuts) >= 0).all()
        if n >= 100:
            work = np.array([np.sum(lens[cuts[r]:cuts[r + 1]] + 1) for r in range(world)])
            assert work.max() <= work.mean() * 1.1 + 301


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, reads, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        compiled = util.compiled_for(["l4c4"], dict(length=4), True)

        def decode(shard):
            out = []
            for r in shard:
                o = util.oracle_viterbi(compiled, r, want_path=False)
                out.append((o["decoded"], util.hexf(o["loglike"])))
            return out

        res = sharding.decode_sharded(decode, reads, rank, world)
        if rank == 0:
            q.put(res)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_decode_matches_single_process(world):
    import torch.multiprocessing as mp
    case = util.golden_case("l4c4_global_mixed")
    reads = [r["seq"] for r in case["reads"]][:12] + ["", "ACGT"]
    want = [(r["decoded"], util.hexf(r["loglike_hex"])) for r in case["reads"][:12]]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, reads, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(got) == len(reads)
    assert got[:12] == want  # input order restored, bit-exact with the reference's golden values
