"""Stress test for the sweep-boundary hazard of chain_kernel: with one launch stage (forced team size or
few chains) a greedy sweep that moves nothing ends with warps >= 1 writing sites / hv after the round's only barrier,
and the next sweep's block loads (sequences 0..63) read them. Small N puts every sequence into those two blocks.
Every run must be bit-identical to the first and to the oracle."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import background_of, planted_motif_set

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("team", [4, 8, 16])
@pytest.mark.parametrize("n", [17, 24, 32, 61])
def test_repeated_runs_are_bit_identical_for_small_sets(n, team):
    L, k = 90, 7
    ps = planted_motif_set(n, L, k, seed=100 + n)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    params = make_params(k, 1e-4, 5, bg)
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    with GibbsEngine(seqs) as eng:
        eng.set_team_warps(team)
        first = eng.run(params, 64, chain_id_base=9, seed=123, want_counts=False)
        for _ in range(25):
            again = eng.run(params, 64, chain_id_base=9, seed=123, want_counts=False)
            assert again.sites.tobytes() == first.sites.tobytes()
            assert again.scores.tobytes() == first.scores.tobytes()
    for c in (0, 31, 63):
        rng, _ = O.make_rng(seed=123, chain=9 + c)
        score, pos, _ = O.site_step("do_site_sampling_with_bpv", S, k, 1e-4, pcv=pcv, rng=rng)
        assert first.sites[c].tolist() == pos.tolist()
        np.testing.assert_allclose(first.scores[c], score, rtol=1e-5)
