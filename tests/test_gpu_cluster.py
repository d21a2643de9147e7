"""The last hand-over stages give one chain a thread-block cluster of 4 / 8 CTAs (chain_cluster_kernel): DSMEM flags, one
cluster barrier per round. A chain's result depends on (seed, chain id) only -- whatever the number of stages, the
cluster size or the number of chains that reach a cluster stage; spot-checked against the oracle."""
import numpy as np
import pytest

import oracle_lib as O
from gibbssampling_b200 import _abi
from gibbssampling_b200.engine import GibbsEngine, make_params
from gibbssampling_b200.synthetic import background_of, planted_motif_set

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(150, 90, None, 8), (260, 200, 120, 12), (129, 70, None, 21)], ids=lambda s: f"n{s[0]}_L{s[1]}_k{s[3]}")
def test_cluster_stages_do_not_change_a_chain(shape):
    n, L, Lmin, k = shape
    ps = planted_motif_set(n, L, k, seed=300 + n, min_length=Lmin)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    params = make_params(k, 1e-4, 5, bg)
    with GibbsEngine(seqs) as eng:
        eng.set_option(_abi.GIBBS_OPT_CLUSTER, 0)
        eng.set_team_warps(4)
        ref = eng.run(params, 700, chain_id_base=3, seed=42, want_counts=False)        # one stage, 4 warps per chain
        eng.set_team_warps(0)
        launches = {}
        for cluster in (0, 4, 8):
            eng.set_option(_abi.GIBBS_OPT_CLUSTER, cluster)
            for chains, base in ((700, 3), (40, 100), (5, 650), (1, 77)):
                res = eng.run(params, chains, chain_id_base=base, seed=42, want_counts=False)
                lo = base - 3
                assert res.sites.tobytes() == ref.sites[lo:lo + chains].tobytes(), (cluster, chains)
                assert res.scores.tobytes() == ref.scores[lo:lo + chains].tobytes(), (cluster, chains)
                assert res.sums.tobytes() == ref.sums[lo:lo + chains].tobytes(), (cluster, chains)
                launches[(cluster, chains)] = res.stats["kernel_launches"]
        assert launches[(4, 700)] == launches[(0, 700)] + 1 and launches[(8, 700)] >= launches[(4, 700)]
    S = O.sources(seqs)
    pcv = O.pcv_from_acgt(bg)
    for c in (0, 350, 699):
        rng, _ = O.make_rng(seed=42, chain=3 + c)
        score, pos, _, _ = O.fast_site_pipeline(0, S, k, 1e-4, pcv=pcv, rng=rng)
        assert ref.sites[c].tolist() == pos.tolist(), f"chain {c}"
        np.testing.assert_allclose(ref.scores[c], score, rtol=1e-5)


def test_phase_functions_through_the_cluster_stage():
    """A caller-supplied start state (NaN marker in hv: the accept test compares log2 scores) and single phases."""
    n, L, k = 160, 80, 7
    ps = planted_motif_set(n, L, k, seed=9)
    seqs = ps.sequences()
    bg = background_of(ps.ascii, 1e-4, 5)
    rng = np.random.default_rng(1)
    sites = rng.integers(0, L - k + 1, size=(6, n)).astype(np.int32)
    scores = rng.normal(size=(6, n)) * 2
    with GibbsEngine(seqs) as eng:
        out = {}
        for cluster in (0, 8):
            eng.set_option(_abi.GIBBS_OPT_CLUSTER, cluster)
            for mask in (_abi.PHASE_GREEDY, _abi.PHASE_LEFT, _abi.PHASE_RIGHT, _abi.PHASE_GREEDY | _abi.PHASE_RIGHT):
                eng.set_start_state(sites, scores)
                r = eng.run(make_params(k, 1e-4, 5, bg, phase_mask=mask), 6, seed=1, want_counts=False)
                out[(cluster, mask)] = (r.sites.tobytes(), r.scores.tobytes())
        for mask in (_abi.PHASE_GREEDY, _abi.PHASE_LEFT, _abi.PHASE_RIGHT, _abi.PHASE_GREEDY | _abi.PHASE_RIGHT):
            assert out[(0, mask)] == out[(8, mask)], mask
