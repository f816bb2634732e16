"""Sharded workloads (BASELINE configs 4 and 5 in small): two ranks, each with its own nid_ctx, own whole problems and
gather one result row per problem through shard.py; the gathered table must equal the single-rank run bit for bit.
Both ranks share cuda:0 here (the round-end GPU test box has one GPU), so the collective runs over gloo; on the 8-GPU
box bench.py runs the same code over NCCL."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problems(synth, orc, n):
    seqs = [synth.make_sequence(3000 + s, 4, 120, 160) for s in range((n + 3) // 4)]
    pairs = [seqs[i // 4][i % 4] for i in range(n)]
    pose0 = np.stack([orc.se3_mul(orc.se3_exp(np.random.default_rng(50 + i).uniform(-1, 1, 6) * 1e-3),
                                  orc.reference_perturbation(p.T_wc1)) for i, p in enumerate(pairs)])
    return pairs, pose0


def _solve_rows(nid, orc, pairs, pose0, idx):
    n = len(idx)
    ctx = nid.Context(120, 160, 4, 16, n_pairs=max(n, 1), max_jobs=max(n, 1))
    sel = [pairs[i] for i in idx]
    keep = ctx.set_pairs_u16(0, np.stack([p.depth0_u16 for p in sel]), np.stack([p.im0 for p in sel]), np.stack([p.im1 for p in sel]),
                             np.stack([p.T_wc0 for p in sel]), np.stack([p.intr for p in sel]))
    ctx.prepare_pairs(0, np.stack([orc.se3_to_mat16(pose0[i]) for i in idx]))
    out, st = ctx.solve_jobs(pose0[idx], np.arange(n, dtype=np.int32), 6)
    return np.concatenate([out, st.astype(np.float64)], axis=1)


def _sweep_rows(nid, orc, synth, p, poses, idx):
    ctx = nid.Context(120, 160, 8, 8, n_pairs=1, max_jobs=max(len(idx), 1))
    ctx.set_pair(0, p.depth0, p.im0, p.im1, p.T_wc0, p.intr)
    tot, _ = ctx.hard_eval_jobs(np.ascontiguousarray(poses[idx]))
    return tot


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nid = importlib.import_module("nid-pose-estimation_b200")
    synth = importlib.import_module("nid-pose-estimation_b200.synth")
    shard = importlib.import_module("nid-pose-estimation_b200.shard")
    from oracle import binding as orc
    pairs, pose0 = _problems(synth, orc, 10)
    mine = shard.owned(10, world, rank, "block")
    t4 = shard.gather_results(_solve_rows(nid, orc, pairs, pose0, mine), 10, world, rank, "block")
    p = pairs[0]
    gt = orc.se3_from_mat16(synth.mat16_inverse(p.T_wc1))
    rng = np.random.default_rng(8)
    poses = np.stack([orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(rng.uniform(-0.03, 0.03, 6)), gt)) for _ in range(37)])
    mine5 = shard.owned(37, world, rank, "cyclic")
    t5 = shard.gather_results(_sweep_rows(nid, orc, synth, p, poses, mine5), 37, world, rank, "cyclic")
    if rank == 0:
        q.put((t4, t5))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_tables_equal_the_single_rank_run(nid, orc, synth):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t4, t5 = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    pairs, pose0 = _problems(synth, orc, 10)
    ref4 = _solve_rows(nid, orc, pairs, pose0, np.arange(10))
    assert np.array_equal(t4, ref4)
    p = pairs[0]
    gt = orc.se3_from_mat16(synth.mat16_inverse(p.T_wc1))
    rng = np.random.default_rng(8)
    poses = np.stack([orc.se3_to_mat16(orc.se3_mul(orc.se3_exp(rng.uniform(-0.03, 0.03, 6)), gt)) for _ in range(37)])
    ref5 = _sweep_rows(nid, orc, synth, p, poses, np.arange(37))
    assert np.array_equal(t5[:, 0], ref5)
    # and the solves are the oracle's: converged poses of two of the problems
    P = orc.Problem(pairs[7].im0, pairs[7].depth0, pairs[7].im1, pairs[7].T_wc0, pairs[7].intr, 4, 16, threads=4)
    P.set_quirks(0, 1)
    P.prepare(pose0[7])
    poseo, its, _, _ = P.optimize(pose0[7], 6)
    assert t4[7, 7] == its and np.max(np.abs(t4[7, :7] - poseo)) < 1e-7
