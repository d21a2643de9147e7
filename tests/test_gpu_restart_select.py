"""The promote-or-restart loop (fs:435-459, quirk A.6-8) decided on the device (gibbs_fetch_best) and, for several
devices in one process, on the host from one sum per restart (gibbs_multi_fetch_best): both must return what the
sequential loop returns -- checked against the host-side model replay_restart_loop (itself checked against the oracle's
or_best_information_content in test_host_logic.py) and against the oracle directly."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import MotifSampler, SiteSampler, _abi
from gibbssampling_b200.CompositeVector import ProbabilityCompositeVector
from gibbssampling_b200.engine import GibbsEngine, MultiEngine, device_count, make_params
from gibbssampling_b200.synthetic import background_of, planted_motif_set

pytestmark = pytest.mark.gpu
DNA = list("ATGC-")


def _case(n=14, L=60, k=7, seed=3):
    ps = planted_motif_set(n, L, k, seed=seed)
    bg = background_of(ps.ascii, 1e-4, 5)
    return ps, bg


@pytest.mark.parametrize("reps", [0, 1, 2, 3, 7, 40, 63])
def test_fetch_best_equals_the_replayed_loop(reps):
    ps, bg = _case()
    params = make_params(7, 1e-4, 5, bg)
    with GibbsEngine(ps.sequences()) as eng:
        res = eng.run(params, reps + 1, chain_id_base=11, seed=5, want_counts=False)
        want = SiteSampler.replay_restart_loop(reps, res.scores, res.sites, res.sums)
        best = eng.fetch_best(reps, want_counts=True)
    got = [(float(s), int(p)) for s, p in zip(best.scores, best.sites)]
    assert got == want
    if best.restart >= 0:
        assert best.total == res.sums[best.restart]
        assert best.sites.tolist() == res.sites[best.restart].tolist()
        # PWM counts of the winner
        want_counts = np.zeros((7, 4), dtype=np.int64)
        for i, s in enumerate(ps.sequences()):
            for j in range(7):
                want_counts[j, "ACGT".index(chr(s[best.sites[i] + j]))] += 1
        assert best.counts.tolist() == want_counts.tolist()
    else:
        assert got == [(0.0, 0)] and reps == 0


def test_fetch_best_matches_the_oracle_restart_loop():
    ps, bg = _case(n=9, L=40, k=5, seed=8)
    S = O.sources(ps.sequences())
    pcv = O.pcv_from_acgt(bg)
    pv = ProbabilityCompositeVector.ofACGT(*bg)
    for reps in (1, 4, 9):
        rng, _ = O.make_rng(seed=77, chain=100)
        score, pos, _ = O.best_information_content(0, reps, S, 5, 1e-4, rng, pcv=pcv)
        got = SiteSampler.getMotifsWithBestInformationContentWithBPV(reps, 5, 1e-4, DNA, ps.sequences(), pv, seed=77, chain=100)
        assert [p for _, p in got] == pos.tolist()
        np.testing.assert_allclose([s for s, _ in got], score, rtol=1e-5)
        rng, _ = O.make_rng(seed=78, chain=5)
        score, pos, _ = O.best_information_content(1, reps, S, 5, 1e-4, rng)
        got = SiteSampler.getMotifsWithBestInformationContent(reps, 5, 1e-4, DNA, ps.sequences(), seed=78, chain=5)
        assert [p for _, p in got] == pos.tolist()
        rng, _ = O.make_rng(seed=79, chain=2)
        want, _ = O.best_motif_information_content(0, reps, S, 1, 5, 1e-4, 0.5, rng, pcv=pcv)
        got = MotifSampler.findBestInormationContentContainingMotifsWithPCV(reps, 1, 5, 1e-4, 0.5, DNA, ps.sequences(), pv,
                                                                            seed=79, chain=2)
        assert [tuple(m.Positions) for m in got] == [tuple(p) for _, p in want]
        np.testing.assert_allclose([m.PWMS for m in got], [v for v, _ in want], rtol=1e-5)


def test_equal_restarts_end_the_loop_and_ties_keep_the_first():
    """Identical restarts (same chain id twice is impossible, so: a set whose restarts all converge to one answer).
    `acc = best` (fs:439) must end the loop exactly where the sequential loop ends."""
    seqs = ["ACGTACGTAC" + "TTGACA" + "GGGCCCGGGC"] * 6     # every restart finds the same sites and scores
    bg = background_of(np.frombuffer("".join(seqs).encode(), dtype=np.uint8), 1.0, 4)
    params = make_params(6, 1.0, 4, bg)
    with GibbsEngine(seqs) as eng:
        res = eng.run(params, 12, seed=1, want_counts=False)
        for reps in (1, 2, 5, 11):
            want = SiteSampler.replay_restart_loop(reps, res.scores, res.sites, res.sums)
            best = eng.fetch_best(reps)
            assert [(float(s), int(p)) for s, p in zip(best.scores, best.sites)] == want


def test_motif_flavour_of_the_initial_value():
    ps, bg = _case(n=8, L=50, k=6, seed=4)
    params = make_params(6, 1e-4, 5, bg, sampler=_abi.GIBBS_MOTIF_SAMPLER, cutoff=2.0)
    with GibbsEngine(ps.sequences()) as eng:
        res = eng.run(params, 6, seed=9, want_counts=False)
        for reps in (0, 1, 5):
            want = MotifSampler.replay_motif_restart_loop(reps, res.scores, res.sites, res.sums)
            best = eng.fetch_best(reps)
            got = [MotifSampler.MotifIndex(float(s), (int(p),) if p >= 0 else ()) for s, p in zip(best.scores, best.sites)]
            assert got == want


@pytest.mark.parametrize("slots", [1, 2, 3])
def test_multi_device_handle_equals_one_device(slots):
    """gibbs_multi_*: restarts split into contiguous blocks over device slots. With fewer GPUs than slots the slots
    share a device (two handles on one GPU) -- the partition, the global chain ids and the host-side loop are the same."""
    ps, bg = _case(n=16, L=70, k=8, seed=6)
    params = make_params(8, 1e-4, 5, bg)
    ndev = device_count()
    devices = [i % ndev for i in range(slots)]
    with GibbsEngine(ps.sequences()) as eng:
        for chains in (1, 5, 37):
            res = eng.run(params, chains, chain_id_base=200, seed=21, want_counts=False)
            want = SiteSampler.replay_restart_loop(chains - 1, res.scores, res.sites, res.sums)
            with MultiEngine(ps.sequences(), devices=devices) as m:
                assert m.n_devices == slots
                m.run_device(params, chains, chain_id_base=200, seed=21)
                best = m.fetch_best(chains - 1, want_counts=True)
            got = [(float(s), int(p)) for s, p in zip(best.scores, best.sites)]
            assert got == want
            assert best.stats["site_updates"] == res.stats["site_updates"]
            if best.restart >= 0:
                assert best.sites.tolist() == res.sites[best.restart].tolist()
                one = eng.fetch_best(chains - 1, want_counts=True)
                assert one.restart == best.restart and one.counts.tolist() == best.counts.tolist()


def test_multi_device_uses_every_gpu_of_the_box():
    ndev = device_count()
    if ndev < 2:
        pytest.skip("one GPU visible")
    ps, bg = _case(n=20, L=80, k=8, seed=7)
    params = make_params(8, 1e-4, 5, bg)
    with MultiEngine(ps.sequences()) as m, GibbsEngine(ps.sequences()) as eng:
        assert m.n_devices == ndev
        m.run_device(params, 4 * ndev + 1, chain_id_base=0, seed=3)
        best = m.fetch_best(4 * ndev)
        res = eng.run(params, 4 * ndev + 1, seed=3, want_counts=False)
        want = SiteSampler.replay_restart_loop(4 * ndev, res.scores, res.sites, res.sums)
    assert [(float(s), int(p)) for s, p in zip(best.scores, best.sites)] == want
