"""The first sweeps of a restart on one warp per chain (GIBBS_OPT_SEQ_SWEEPS): when many chains share the GPU, the early
greedy sweeps (almost every update moves a site, so the in-place sweep of fs:388 is sequential) run as stages of their own
on chain_kernel<KP, 1>; every chain pauses at its first sweep boundary and the next stage continues it from the stored
(phase, sweeps in phase). The hand-over must change nothing: same sites, scores, sums and sweep counts for 0, 1, 2 or 3
such sweeps, equal to the oracle; also when a restart converges inside those stages, when max_sweeps cuts the phase short
and when the run has no greedy phase at all."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import background_of, planted_motif_set

pytestmark = pytest.mark.gpu


def _many_chains():
    """enough chains that the first team stage is the four-warp one (more than GIBBS_OPT_STAGE2_AT = 4 chains per SM):
    148 SMs on a B200"""
    return 4 * 148 + 40


@pytest.mark.parametrize("n,L,Lmin,k", [(70, 120, 90, 9), (64, 300, None, 12), (130, 90, None, 20)])
def test_one_warp_stages_change_nothing(n, L, Lmin, k):
    ps = planted_motif_set(n, L, k, seed=300 + n, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    params = make_params(k, 1e-4, 5, bg)
    chains = _many_chains()
    runs = {}
    with GibbsEngine(seqs) as eng:
        eng.run(params, chains, chain_id_base=11, seed=1, want_counts=False)   # (the first run also builds the W table: one launch more)
        for s in (0, 1, 2, 3):
            eng.set_option(_abi.GIBBS_OPT_SEQ_SWEEPS, s)
            runs[s] = eng.run(params, chains, chain_id_base=11, seed=4242, want_counts=False)
    base = runs[0]
    for s in (1, 2, 3):
        r = runs[s]
        assert r.sites.tobytes() == base.sites.tobytes(), s
        assert r.scores.tobytes() == base.scores.tobytes(), s
        assert r.sums.tobytes() == base.sums.tobytes(), s
        assert r.stats["sweeps"] == base.stats["sweeps"] and r.stats["site_updates"] == base.stats["site_updates"], s
        assert r.stats["kernel_launches"] == base.stats["kernel_launches"] + 1, s   # the one-warp stage did run
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    scores, pos, sums, _ = O.fast_site_chains(0, S, k, 1e-4, seed=4242, chain_base=11, n_chains=24, pcv=pcv, threads=4)
    assert runs[2].sites[:24].tolist() == pos.tolist()
    np.testing.assert_allclose(runs[2].scores[:24], scores, rtol=1e-5)


def test_one_warp_stages_with_capped_sweeps_and_phase_masks():
    n, L, k = 80, 100, 8
    ps = planted_motif_set(n, L, k, seed=77)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    chains = _many_chains()
    with GibbsEngine(seqs) as eng:
        for kw in (dict(max_sweeps=1), dict(max_sweeps=2), dict(phase_mask=_abi.PHASE_INIT | _abi.PHASE_GREEDY),
                   dict(phase_mask=_abi.PHASE_INIT | _abi.PHASE_LEFT | _abi.PHASE_RIGHT), dict(phase_shifts=False)):
            params = make_params(k, 1e-4, 5, bg, **kw)
            eng.set_option(_abi.GIBBS_OPT_SEQ_SWEEPS, 0)
            want = eng.run(params, chains, seed=5, want_counts=False)
            eng.set_option(_abi.GIBBS_OPT_SEQ_SWEEPS, 2)
            got = eng.run(params, chains, seed=5, want_counts=False)
            assert got.sites.tobytes() == want.sites.tobytes(), kw
            assert got.scores.tobytes() == want.scores.tobytes(), kw
            assert got.stats["sweeps"] == want.stats["sweeps"], kw
            assert got.stats["capped_chains"] == want.stats["capped_chains"], kw
