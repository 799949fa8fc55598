"""N>1 host logic on CPU: world_size-2 gloo processes shard the frequency grid / the profile ensemble,
compute their block (with the oracle standing in for the GPU kernels -- tests may do that), gather with
torch.distributed and must reproduce the single-process result exactly.  On a multi-GPU box the same
logic runs with NCCL through bench.py."""
import os
import socket
import sys

import numpy as np
import pytest

from monortm_b200 import sharding

HERE = os.path.dirname(os.path.abspath(__file__))


def test_block_partition_covers_everything():
    for n in (1, 2, 7, 100, 125000, 1000003):
        for world in (1, 2, 3, 4, 8):
            blocks = [sharding.block_partition(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_freq_shard_keeps_global_range():
    wn = np.linspace(0.2, 1.2, 101)
    parts = [sharding.freq_shard(wn, r, 3) for r in range(3)]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), wn)
    for r, (w, (v1, v2, iw0)) in enumerate(parts):
        assert v1 == wn[0] and v2 == wn[-1] and wn[iw0] == w[0]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist
    import harness
    from monortm_b200 import synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- frequency sharding (configs 3, 5): gridded continuum mode needs the global origin + offset
        wn, dv = synth.freq_c1_sweep()
        case = harness.make_case(n_filler=128, nlay=5, wn=wn, dvset=dv, irt=1)
        wl, (v1, v2, iw0) = sharding.freq_shard(wn, rank, world)
        pr = case["prof"]
        # the oracle has no shard options: emulate the global origin by list-mode evaluation of the same grid
        m = harness.oracle_modm(case["ls"], wl, 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22,
                                pr["wkl"][:, :, 0], pr["wbrodl"][:, 0], case["scor"][:, :, :, 0])
        r = harness.oracle_rtm(1, 1, wl, pr["t"][:, 0], pr["tz"][:, 0], m["o"], 290.0, case["reflc"][iw0:iw0 + len(wl)],
                               case["emiss"][iw0:iw0 + len(wl)])
        local = torch.from_numpy(np.stack([r["rad"], r["tb"], r["trtot"]], axis=1))
        counts = [sharding.block_partition(len(wn), k, world)[1] for k in range(world)]
        full = sharding.gather_blocks(local, counts)
        # --- profile sharding (config 4)
        prof = synth.synthetic_profiles(5, 4, seed0=77)
        mine, start = sharding.profile_shard(prof, rank, world)
        tsum = torch.from_numpy(np.ascontiguousarray(mine["t"].sum(axis=0)))
        pc = [sharding.block_partition(5, k, world)[1] for k in range(world)]
        tall = sharding.gather_blocks(tsum, pc)
        if rank == 0:
            q.put((full.numpy(), tall.numpy()))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharded_equals_single_process():
    import torch.multiprocessing as mp
    import harness
    from monortm_b200 import synth
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full, tall = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    wn, dv = synth.freq_c1_sweep()
    case = harness.make_case(n_filler=128, nlay=5, wn=wn, dvset=dv, irt=1)
    pr = case["prof"]
    m = harness.oracle_modm(case["ls"], wn, 0.0, pr["p"][:, 0], pr["t"][:, 0], pr["clw"][:, 0], 22,
                            pr["wkl"][:, :, 0], pr["wbrodl"][:, 0], case["scor"][:, :, :, 0])
    r = harness.oracle_rtm(1, 1, wn, pr["t"][:, 0], pr["tz"][:, 0], m["o"], 290.0, case["reflc"], case["emiss"])
    assert np.array_equal(full[:, 0], r["rad"]) and np.array_equal(full[:, 1], r["tb"]) and np.array_equal(full[:, 2], r["trtot"])
    prof = synth.synthetic_profiles(5, 4, seed0=77)
    assert np.array_equal(tall, prof["t"].sum(axis=0))
