"""Independent pure-Python model of the reference (TEST INFRASTRUCTURE ONLY, small inputs).

Written directly from /root/reference/GibbsSampling/GibbsSampling.fs in a different style from the C
oracle (dicts keyed by symbol, lists of tuples) so the two restatements can cross-check each other.
Python floats are IEEE float64 evaluated left to right like the F#.
"""
from __future__ import annotations

import math


def log2(x: float) -> float:  # FSharpAux: Math.Log(x, 2.0)
    if x == 0.0:
        return -math.inf
    return math.log(x) / math.log(2.0)


# ---- CompositeVector ---------------------------------------------------------------------------
def fcv_of(seq):  # fs:60
    v = {}
    for s in seq:
        v[s] = v.get(s, 0) + 1
    return v


def fcv_without(k, pos, seq):  # fs:73
    return fcv_of(seq[:pos] + seq[pos + k:])


def fuse_fcv(alphabet, vecs):  # fs:65 (alphabet slots only)
    out = {}
    for v in vecs:
        for a in alphabet:
            out[a] = out.get(a, 0) + v.get(a, 0)
    return out


def normalized_pcv(alphabet, pc, fcv):  # fs:115
    total = float(sum(fcv.values())) + (float(len(alphabet)) * pc)
    pcv = {s: float(c) for s, c in fcv.items()}
    for a in alphabet:
        pcv[a] = (pcv.get(a, 0.0) + pc) / total
    return pcv


def pcv_score(pcv, seg):  # fs:123
    v = 1.0
    for s in seg:
        v = v * pcv.get(s, 0.0)
    return v


# ---- PositionMatrix ------------------------------------------------------------------------------
def loo_counts(seqs, k, positions, heldout):
    """fs:392-396 with positions[i] a list of sites (MotifSampler) or an int (SiteSampler)."""
    cnt = {}
    for i, s in enumerate(seqs):
        if i == heldout:
            continue
        ps = positions[i] if isinstance(positions[i], (list, tuple)) else [positions[i]]
        for p in ps:
            for j in range(k):
                cnt[(s[p + j], j)] = cnt.get((s[p + j], j), 0) + 1
    return cnt


def ppm_of(cnt, k, source_count, alphabet, pc):  # fs:249-261
    den = float(source_count) + (float(len(alphabet)) * pc)
    ppm = {key: float(c) for key, c in cnt.items()}
    for a in alphabet:
        for j in range(k):
            ppm[(a, j)] = (ppm.get((a, j), 0.0) + pc) / den
    return ppm


def pwm_of(alphabet, pcv, ppm, k):  # fs:282
    return {(a, j): ppm[(a, j)] / pcv[a] for a in alphabet for j in range(k)}


def pwm_score(pwm, seg):  # fs:290
    v = 1.0
    for j, s in enumerate(seg):
        v = v * pwm.get((s, j), 0.0)
    return v


# ---- SiteSampler -----------------------------------------------------------------------------------
def best_with_bpv(k, alphabet, seq, pcv, ppm):  # fs:301
    hv, hi = 0.0, 0
    for n in range(len(seq) - k + 1):
        tmp = pwm_score(pwm_of(alphabet, pcv, ppm, k), seq[n:n + k])
        if tmp > hv:
            hv, hi = tmp, n
    return log2(hv), hi


def best_drifting(k, alphabet, pc, seq, fcv, ppm):  # fs:462 (fcv mutated in place)
    hv, hi = 0.0, 0
    for n in range(len(seq) - k + 1):
        seg = seq[n:n + k]
        for s in seq:
            fcv[s] = fcv.get(s, 0) + 1
        for s in seg:
            fcv[s] = fcv[s] - 1 if fcv[s] - 1 > 0 else 0
        pcv = normalized_pcv(alphabet, pc, fcv)
        tmp = pwm_score(pwm_of(alphabet, pcv, ppm, k), seg)
        if tmp > hv:
            hv, hi = tmp, n
    return log2(hv), hi


def _shift(mode, pos, length, k):
    if mode == "left":
        return pos - 1 if pos > 0 else pos
    if mode == "right":
        return pos + 1 if pos <= length - k - 1 else pos
    return pos


def _scan(k, pc, alphabet, seqs, pcv, positions, h):
    cnt = loo_counts(seqs, k, positions, h)
    ppm = ppm_of(cnt, k, len(seqs) - 1, alphabet, pc)
    if pcv is not None:
        return best_with_bpv(k, alphabet, seqs[h], pcv, ppm)
    fcv = fuse_fcv(alphabet, [fcv_without(k, positions[i], seqs[i]) for i in range(len(seqs)) if i != h])
    return best_drifting(k, alphabet, pc, seqs[h], fcv, ppm)


def sweep_until_stable(k, pc, alphabet, seqs, pcv, start, mode="none"):  # fs:381 / 350 / 318 / 554 / 519 / 483
    acc = list(start)
    snap = list(start)
    while True:
        for h in range(len(seqs)):
            src = acc if mode == "none" else snap
            positions = [_shift(mode, p, len(seqs[i]), k) for i, (_, p) in enumerate(src)]
            tmp = _scan(k, pc, alphabet, seqs, pcv, positions, h)
            if tmp[0] > acc[h][0]:
                acc[h] = tmp
        if [p for _, p in acc] == [p for _, p in snap]:
            return acc
        snap = list(acc)


def random_starts(k, pc, alphabet, seqs, pcv, draws):  # fs:412 / fs:589; draws = iterator of uniforms
    out = []
    for h in range(len(seqs)):
        positions = [0] * len(seqs)
        for i in range(len(seqs)):
            if i != h:
                positions[i] = int(next(draws) * float(len(seqs[i]) - k + 1))
        out.append(_scan(k, pc, alphabet, seqs, pcv, positions, h))
    return out


def do_site_sampling(k, pc, alphabet, seqs, pcv, draws):  # fs:691 (pcv given) / fs:697 (pcv None)
    st = random_starts(k, pc, alphabet, seqs, pcv, draws)
    st = sweep_until_stable(k, pc, alphabet, seqs, pcv, st, "none")
    st = sweep_until_stable(k, pc, alphabet, seqs, pcv, st, "left")
    return sweep_until_stable(k, pc, alphabet, seqs, pcv, st, "right")


def restart_loop(reps, run_restart):  # fs:435-459; run_restart() -> list of (score, pos)
    n, acc, best = 0, [], [(0.0, 0)]
    while True:
        if n > reps or acc == best:
            return best
        ia = 0.0
        for s, _ in acc:
            ia = ia + s
        ib = 0.0
        for s, _ in best:
            ib = ib + s
        if ia > ib:
            n, acc, best = n + 1, [], (acc if acc else best)
        else:
            n, acc = n + 1, run_restart()


# ---- MotifSampler ----------------------------------------------------------------------------------
def check_distance(width, items):  # fs:129
    return all(abs(items[a] - items[b]) > width for a in range(len(items)) for b in range(a + 1, len(items)))


def combos(cutoff, width, m, scored):  # fs:727
    out = []

    def loop(prob, positions, size, rest):
        if rest:
            x, xs = rest[0], rest[1:]
            if size > 0 and check_distance(width, [x[1]] + positions) and log2(x[0] * prob) > cutoff:
                loop(x[0] * prob, [x[1]] + positions, size - 1, xs)
            if size >= 0:
                loop(prob, positions, size, xs)
        elif size == 0:
            out.append((log2(prob), positions))

    loop(1.0, [], m, scored)
    return out


def candidate_list(cutoff, m, k, seq, pcv, pwm):  # fs:759
    segs = [(seq[n:n + k], n) for n in range(len(seq) - k + 1)]
    scored = [(pwm_score(pwm, s), n) for s, n in segs]
    items = [(pcv_score(pcv, s), []) for s, _ in segs]
    for size in range(1, m + 1):
        items += combos(cutoff, k, size, scored)
    return items


def roulette(pick, items):  # fs:746
    total = 0.0
    for w, _ in items:
        total = total + w
    acc = 0.0
    for idx, (w, _) in enumerate(items):
        wn = w / total
        if acc <= pick and pick <= acc + wn:
            return idx
        acc = acc + wn
    raise IndexError("pick beyond the accumulated mass (fs:753)")


def _motif_tables(k, pc, alphabet, seqs, pcv_fixed, state, h):
    cnt = loo_counts(seqs, k, [p for _, p in state], h)
    ppm = ppm_of(cnt, k, len(seqs) - 1, alphabet, pc)
    if pcv_fixed is not None:
        pcv = pcv_fixed
    else:  # fs:896-905
        vecs = [fcv_without(k, p, seqs[i]) for i, (_, ps) in enumerate(state) if i != h for p in ps]
        fcv = fuse_fcv(alphabet, vecs)
        for s in seqs[h]:
            fcv[s] = fcv.get(s, 0) + 1
        pcv = normalized_pcv(alphabet, pc, fcv)
    return pcv, pwm_of(alphabet, pcv, ppm, k)


def motif_stochastic(m, k, pc, cutoff, alphabet, seqs, pcv_fixed, state, draws):  # fs:828 / fs:935
    out = []
    for h in range(len(seqs)):
        pcv, pwm = _motif_tables(k, pc, alphabet, seqs, pcv_fixed, state, h)
        items = candidate_list(cutoff, m, k, seqs[h], pcv, pwm)
        out.append(items[roulette(next(draws), items)])
    return out


def motif_greedy(m, k, pc, cutoff, alphabet, seqs, pcv_fixed, state):  # fs:788 / fs:885
    acc, snap = list(state), list(state)
    while True:
        for h in range(len(seqs)):
            pcv, pwm = _motif_tables(k, pc, alphabet, seqs, pcv_fixed, acc, h)
            items = candidate_list(cutoff, m, k, seqs[h], pcv, pwm)
            best = items[0]
            for it in items[1:]:
                if it[0] > best[0]:
                    best = it
            if best[0] > acc[h][0]:
                acc[h] = best
        if [p for _, p in acc] == [p for _, p in snap]:
            return acc
        snap = list(acc)


def do_motif_sampling(m, k, pc, cutoff, alphabet, seqs, pcv_fixed, draws):  # fs:876-879 / fs:1034
    st = [(s, [p]) for s, p in random_starts(k, pc, alphabet, seqs, pcv_fixed, draws)]
    st = motif_stochastic(m, k, pc, cutoff, alphabet, seqs, pcv_fixed, st, draws)
    return motif_greedy(m, k, pc, cutoff, alphabet, seqs, pcv_fixed, st)
