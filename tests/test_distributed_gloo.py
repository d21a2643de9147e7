"""The N > 1 path on CPU: world_size-2 gloo process group, chain sharding + the single all_gather."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import oracle_lib as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_chains, seqs, k, pc, out_dir):
    import torch.distributed as dist
    from gibbssampling_b200.distributed import allgather_best, select_best, shard_chains

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        S = O.sources(seqs)
        pcv = O.pcv_of_sources(S, pc)
        first, count = shard_chains(n_chains, rank, world)
        # the oracle stands in for the GPU kernel: this test covers the host-side sharding and exchange only
        sums, ids, sites, scores = [], [], [], []
        for c in range(first, first + count):
            rng, _ = O.make_rng(seed=11, chain=c)
            sc, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, pc, pcv=pcv, rng=rng)
            acc = 0.0
            for v in sc:
                acc = acc + float(v)
            sums.append(acc)
            ids.append(c)
            sites.append(pos)
            scores.append(sc)
        if count:
            b = select_best(np.array(sums), np.array(ids))
            got = allgather_best(sums[b], ids[b], sites[b], scores[b])
        else:
            got = allgather_best(float("-inf"), -1, np.zeros(S.n, np.int32), np.zeros(S.n))
        np.save(os.path.join(out_dir, f"rank{rank}.npy"),
                np.concatenate([[got[0], got[1], got[4]], got[2].astype(np.float64), got[3]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chains", [5, 1])
def test_two_rank_gloo_allgather_picks_the_global_best(tmp_path, golden, n_chains):
    seqs, k, pc = golden["sequences"], golden["k"], golden["pc"]
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_chains, seqs, k, pc, str(tmp_path)), nprocs=world, join=True)
    # single-process answer
    S = O.sources(seqs)
    pcv = O.pcv_of_sources(S, pc)
    best = None
    for c in range(n_chains):
        rng, _ = O.make_rng(seed=11, chain=c)
        sc, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, pc, pcv=pcv, rng=rng)
        acc = 0.0
        for v in sc:
            acc = acc + float(v)
        if best is None or acc > best[0]:
            best = (acc, c, pos, sc)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert r0.tobytes() == r1.tobytes()            # every rank holds the same winner
    n = S.n
    assert r0[0] == best[0] and int(r0[1]) == best[1]
    assert r0[3:3 + n].astype(np.int32).tolist() == best[2].tolist()
    assert r0[3 + n:].tolist() == best[3].tolist()
