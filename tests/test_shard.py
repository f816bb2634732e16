"""Host-side sharding logic: whole problems per rank, one all-gather of results.
CPU only: world_size-2 gloo processes stand in for the per-GPU ranks."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _shard():
    return importlib.import_module("nid-pose-estimation_b200.shard")


@pytest.mark.parametrize("scheme", ["cyclic", "block"])
@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (7, 2), (1024, 8), (24576, 8), (5, 8)])
def test_partition_is_exact_cover(n, world, scheme):
    sh = _shard()
    seen = np.concatenate([sh.owned(n, world, r, scheme) for r in range(world)]) if world else np.zeros(0)
    assert sorted(seen.tolist()) == list(range(n))
    sizes = [sh.owned(n, world, r, scheme).size for r in range(world)]
    assert max(sizes) - min(sizes) <= 1
    assert max(sizes) == (sh.max_owned(n, world) if n else 0)


def test_partition_rejects_bad_arguments():
    sh = _shard()
    with pytest.raises(ValueError):
        sh.owned(4, 2, 2)
    with pytest.raises(ValueError):
        sh.owned(4, 0, 0)
    with pytest.raises(ValueError):
        sh.owned(4, 2, 0, "zigzag")


def test_single_rank_gather_is_identity():
    sh = _shard()
    rows = np.arange(12, dtype=np.float64).reshape(4, 3)
    np.testing.assert_array_equal(sh.gather_results(rows, 4, 1, 0), rows)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, scheme, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = _shard()

    def run_local(idx):
        # stand-in for nid_solve_jobs on this rank's pairs: a deterministic row per problem
        return np.stack([idx * 1.5, idx % 7, np.full(idx.shape, rank)], axis=1).astype(np.float64)

    table = sh.run_sharded(n, world, rank, run_local, scheme)
    q.put((rank, table))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("scheme,n", [("cyclic", 11), ("block", 11), ("cyclic", 1)])
def test_two_rank_gloo_gather_orders_results_by_problem(scheme, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, scheme, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sh = _shard()
    idx = np.arange(n)
    owner = np.zeros(n)
    for r in range(2):
        owner[sh.owned(n, 2, r, scheme)] = r
    want = np.stack([idx * 1.5, idx % 7, owner], axis=1)
    for r in range(2):
        np.testing.assert_array_equal(got[r], want)
