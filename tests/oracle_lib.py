"""ctypes binding of the CPU oracle (oracle/libgibbs_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product package (gibbssampling_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libgibbs_oracle.so")

NSLOT = 49
MAX_M = 4
DNA_BASES = b"ATGC-"  # fsx:368-369: [|A; T; G; C; Gap|]

OK, ERR_ARG, ERR_SYMBOL, ERR_SHORT_SEQ, ERR_ROULETTE = 0, 1, 2, 3, 4


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"oracle error code {code}")
        self.code = code


class Rng(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("u", C.POINTER(C.c_double)),
        ("n_u", C.c_int64),
        ("next", C.c_int64),
        ("seed", C.c_uint64),
        ("chain", C.c_uint64),
        ("exhausted", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("site_updates", C.c_int64),
        ("window_scores", C.c_int64),
        ("sweeps", C.c_int64),
        ("restarts", C.c_int64),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "gibbs_oracle.c")
    hdr = os.path.join(ORACLE_DIR, "gibbs_oracle.h")
    stale = (not os.path.exists(ORACLE_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(ORACLE_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "-B"], check=True, capture_output=True)
    return ORACLE_SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.or_uniform_at.restype = C.c_double
        _lib.or_uniform_at.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        _lib.or_next_uniform.restype = C.c_double
        _lib.or_draw_to_position.restype = C.c_int32
        _lib.or_draw_to_position.argtypes = [C.c_double, C.c_int32, C.c_int32]
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise OracleError(rc)


@dataclass
class Sources:
    """Concatenated ASCII symbols + offsets, the oracle's view of BioArray<_>[]."""

    buf: np.ndarray  # uint8
    off: np.ndarray  # int64 [n+1]

    @property
    def n(self) -> int:
        return len(self.off) - 1

    def length(self, i: int) -> int:
        return int(self.off[i + 1] - self.off[i])

    def seq(self, i: int) -> bytes:
        return self.buf[self.off[i]: self.off[i + 1]].tobytes()


def sources(seqs) -> Sources:
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if off[-1] else np.zeros(0, np.uint8)
    return Sources(buf, off)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(b: bytes):
    return (C.c_uint8 * len(b)).from_buffer_copy(b)


def make_rng(uniforms=None, seed: int = 0, chain: int = 0):
    """Injected stream (uniforms given) or Philox (seed, chain). Returns (Rng, keepalive)."""
    r = Rng()
    keep = None
    if uniforms is not None:
        keep = np.ascontiguousarray(uniforms, dtype=np.float64)
        r.mode = 0
        r.u = _p(keep, C.c_double)
        r.n_u = len(keep)
    else:
        r.mode = 1
        r.seed = seed
        r.chain = chain
    return r, keep


def philox4x32_10(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().or_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def uniform_at(seed: int, chain: int, draw: int) -> float:
    return float(lib().or_uniform_at(seed, chain, draw))


def draw_to_position(u: float, length: int, k: int) -> int:
    return int(lib().or_draw_to_position(u, length, k))


# ---------------------------------------------------------------------------------------------
# primitives
# ---------------------------------------------------------------------------------------------
def loo_pfm(src: Sources, sites, heldout: int, k: int) -> np.ndarray:
    sites = np.ascontiguousarray(sites, dtype=np.int32)
    out = np.zeros((NSLOT, k), dtype=np.int32)
    _check(lib().or_loo_pfm(_p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(src.n),
                            _p(sites, C.c_int32), C.c_int32(heldout), C.c_int32(k), _p(out, C.c_int32)))
    return out


def acgt_counts(pfm49: np.ndarray) -> np.ndarray:
    """[k,4] counts in A,C,G,T order from a 49 x k PFM."""
    rows = [ord(ch) - 42 for ch in "ACGT"]
    return np.ascontiguousarray(pfm49[rows, :].T)


def ppm_of_pfm(pfm49: np.ndarray, source_count: int, pc: float, alphabet: bytes = DNA_BASES) -> np.ndarray:
    k = pfm49.shape[1]
    pfm49 = np.ascontiguousarray(pfm49, dtype=np.int32)
    out = np.zeros((NSLOT, k), dtype=np.float64)
    _check(lib().or_ppm_of_pfm(_p(pfm49, C.c_int32), C.c_int32(k), C.c_int32(source_count), _u8(alphabet),
                               C.c_int32(len(alphabet)), C.c_double(pc), _p(out, C.c_double)))
    return out


def pcv_of_sources(src: Sources, pc: float, alphabet: bytes = DNA_BASES) -> np.ndarray:
    out = np.zeros(NSLOT, dtype=np.float64)
    _check(lib().or_pcv_of_sources(_p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(src.n),
                                   _u8(alphabet), C.c_int32(len(alphabet)), C.c_double(pc), _p(out, C.c_double)))
    return out


def pcv_from_acgt(bg4) -> np.ndarray:
    """49-slot pcv with the given A,C,G,T probabilities (other slots 0)."""
    out = np.zeros(NSLOT, dtype=np.float64)
    for ch, v in zip("ACGT", bg4):
        out[ord(ch) - 42] = v
    return out


def window_scores_bpv(seq: bytes, k: int, pcv: np.ndarray, ppm: np.ndarray, alphabet: bytes = DNA_BASES) -> np.ndarray:
    w = len(seq) - k + 1
    out = np.zeros(max(w, 0), dtype=np.float64)
    pcv = np.ascontiguousarray(pcv, np.float64)
    ppm = np.ascontiguousarray(ppm, np.float64)
    _check(lib().or_window_scores_bpv(_u8(seq), C.c_int32(len(seq)), C.c_int32(k), _u8(alphabet),
                                      C.c_int32(len(alphabet)), _p(pcv, C.c_double), _p(ppm, C.c_double),
                                      _p(out, C.c_double)))
    return out


def best_pwms_with_bpv(seq: bytes, k: int, pcv, ppm, alphabet: bytes = DNA_BASES):
    pcv = np.ascontiguousarray(pcv, np.float64)
    ppm = np.ascontiguousarray(ppm, np.float64)
    s = C.c_double()
    p = C.c_int32()
    _check(lib().or_best_pwms_with_bpv(_u8(seq), C.c_int32(len(seq)), C.c_int32(k), _u8(alphabet),
                                       C.c_int32(len(alphabet)), _p(pcv, C.c_double), _p(ppm, C.c_double),
                                       C.byref(s), C.byref(p)))
    return s.value, p.value


def loo_fcv(src: Sources, sites, heldout: int, k: int, alphabet: bytes = DNA_BASES) -> np.ndarray:
    sites = np.ascontiguousarray(sites, dtype=np.int32)
    out = np.zeros(NSLOT, dtype=np.int32)
    _check(lib().or_loo_fcv(_p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(src.n),
                            _p(sites, C.c_int32), C.c_int32(heldout), C.c_int32(k), _u8(alphabet),
                            C.c_int32(len(alphabet)), _p(out, C.c_int32)))
    return out


def best_pwms(seq: bytes, k: int, pc: float, fcv, ppm, alphabet: bytes = DNA_BASES):
    """fs:462 drifting-background scan. Returns (score, pos, raw window scores, mutated fcv)."""
    fcv = np.ascontiguousarray(fcv, np.int32).copy()
    ppm = np.ascontiguousarray(ppm, np.float64)
    w = len(seq) - k + 1
    raw = np.zeros(max(w, 0), dtype=np.float64)
    s = C.c_double()
    p = C.c_int32()
    _check(lib().or_best_pwms(_u8(seq), C.c_int32(len(seq)), C.c_int32(k), _u8(alphabet), C.c_int32(len(alphabet)),
                              C.c_double(pc), _p(fcv, C.c_int32), _p(ppm, C.c_double), C.byref(s), C.byref(p),
                              _p(raw, C.c_double)))
    return s.value, p.value, raw, fcv


def candidates(seq: bytes, k: int, m: int, cutoff: float, pcv, ppm, alphabet: bytes = DNA_BASES, cap: int = 1 << 20):
    pcv = np.ascontiguousarray(pcv, np.float64)
    ppm = np.ascontiguousarray(ppm, np.float64)
    pw = np.zeros(cap, np.float64)
    npos = np.zeros(cap, np.int32)
    pos = np.zeros(cap * MAX_M, np.int32)
    n = C.c_int64()
    _check(lib().or_candidates(_u8(seq), C.c_int32(len(seq)), C.c_int32(k), C.c_int32(m), C.c_double(cutoff),
                               _u8(alphabet), C.c_int32(len(alphabet)), _p(pcv, C.c_double), _p(ppm, C.c_double),
                               _p(pw, C.c_double), _p(npos, C.c_int32), _p(pos, C.c_int32), C.c_int64(cap),
                               C.byref(n)))
    nn = min(n.value, cap)
    return [(float(pw[i]), [int(x) for x in pos[i * MAX_M: i * MAX_M + npos[i]]]) for i in range(nn)]


def roulette(pwms, pick: float) -> int:
    pwms = np.ascontiguousarray(pwms, np.float64)
    idx = C.c_int64()
    _check(lib().or_roulette(_p(pwms, C.c_double), C.c_int64(len(pwms)), C.c_double(pick), C.byref(idx)))
    return idx.value


# ---------------------------------------------------------------------------------------------
# SiteSampler
# ---------------------------------------------------------------------------------------------
def _site_call(fn_name, src: Sources, k, pc, alphabet, extra_in, rng, score, pos, stats):
    args = [_p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(src.n), C.c_int32(k), C.c_double(pc),
            _u8(alphabet), C.c_int32(len(alphabet))]
    args += extra_in
    if rng is not None:
        args.append(C.byref(rng))
    args += [_p(score, C.c_double), _p(pos, C.c_int32), C.byref(stats)]
    _check(getattr(lib(), fn_name)(*args))


def site_step(name: str, src: Sources, k: int, pc: float, pcv=None, ppm=None, state=None, rng=None,
              alphabet: bytes = DNA_BASES):
    """Run one reference-level function of the SiteSampler.

    name in {random_starts_with_bpv, find_best_motif_with_start_position, left_shifted_with_bpv,
             right_shifted_with_bpv, do_site_sampling_with_bpv, random_starts,
             best_pwms_with_start_positions, left_shifted, right_shifted, do_site_sampling,
             motifs_with_best_pwms_of_ppm, do_site_sampling_with_ppm}
    state = (scores, positions) for the sweep functions. Returns (scores, positions, Stats).
    """
    n = src.n
    if state is not None:
        score = np.ascontiguousarray(state[0], np.float64).copy()
        pos = np.ascontiguousarray(state[1], np.int32).copy()
    else:
        score = np.zeros(n, np.float64)
        pos = np.zeros(n, np.int32)
    st = Stats()
    extra = []
    keep = []
    if name.endswith("with_bpv") or name == "find_best_motif_with_start_position":
        a = np.ascontiguousarray(pcv, np.float64)
        keep.append(a)
        extra.append(_p(a, C.c_double))
    if name.endswith("of_ppm") or name.endswith("with_ppm"):
        a = np.ascontiguousarray(ppm, np.float64)
        keep.append(a)
        extra.append(_p(a, C.c_double))
    needs_rng = name in ("random_starts_with_bpv", "do_site_sampling_with_bpv", "random_starts",
                         "do_site_sampling", "motifs_with_best_pwms_of_ppm", "do_site_sampling_with_ppm")
    _site_call("or_" + name, src, k, pc, alphabet, extra, rng if needs_rng else None, score, pos, st)
    return score, pos, st


def best_information_content(variant: int, reps: int, src: Sources, k: int, pc: float, rng, pcv=None, ppm=None,
                             alphabet: bytes = DNA_BASES):
    n = src.n
    score = np.zeros(max(n, 1), np.float64)
    pos = np.zeros(max(n, 1), np.int32)
    n_out = C.c_int32()
    st = Stats()
    pcv_a = np.ascontiguousarray(pcv, np.float64) if pcv is not None else None
    ppm_a = np.ascontiguousarray(ppm, np.float64) if ppm is not None else None
    _check(lib().or_best_information_content(
        C.c_int32(variant), C.c_int32(reps), _p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(n),
        C.c_int32(k), C.c_double(pc), _u8(alphabet), C.c_int32(len(alphabet)),
        _p(pcv_a, C.c_double) if pcv_a is not None else None,
        _p(ppm_a, C.c_double) if ppm_a is not None else None,
        C.byref(rng), _p(score, C.c_double), _p(pos, C.c_int32), C.byref(n_out), C.byref(st)))
    return score[: n_out.value].copy(), pos[: n_out.value].copy(), st


# ---------------------------------------------------------------------------------------------
# MotifSampler
# ---------------------------------------------------------------------------------------------
def _motif_state(n, state):
    pw = np.zeros(max(n, 1), np.float64)
    npos = np.zeros(max(n, 1), np.int32)
    pos = np.full(max(n, 1) * MAX_M, -1, np.int32)
    if state is not None:
        for i, (v, pl) in enumerate(state):
            pw[i] = v
            npos[i] = len(pl)
            pos[i * MAX_M: i * MAX_M + len(pl)] = pl
    return pw, npos, pos


def _motif_out(pw, npos, pos, n):
    return [(float(pw[i]), [int(x) for x in pos[i * MAX_M: i * MAX_M + npos[i]]]) for i in range(n)]


def motif_step(name: str, variant: int, src: Sources, m: int, k: int, pc: float, cutoff: float, pcv=None,
               ppm=None, state=None, rng=None, alphabet: bytes = DNA_BASES):
    """name in {greedy, stochastic, do_motif_sampling}. state: list of (pwms, [positions])."""
    n = src.n
    pw, npos, pos = _motif_state(n, state)
    st = Stats()
    pcv_a = np.ascontiguousarray(pcv, np.float64) if pcv is not None else None
    ppm_a = np.ascontiguousarray(ppm, np.float64) if ppm is not None else None
    head = [C.c_int32(variant), _p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(n), C.c_int32(m),
            C.c_int32(k), C.c_double(pc), C.c_double(cutoff), _u8(alphabet), C.c_int32(len(alphabet)),
            _p(pcv_a, C.c_double) if pcv_a is not None else None]
    tail = [_p(pw, C.c_double), _p(npos, C.c_int32), _p(pos, C.c_int32), C.byref(st)]
    if name == "greedy":
        _check(lib().or_motif_greedy(*head, *tail))
    elif name == "stochastic":
        _check(lib().or_motif_stochastic(*head, C.byref(rng), *tail))
    elif name == "do_motif_sampling":
        _check(lib().or_do_motif_sampling(*head, _p(ppm_a, C.c_double) if ppm_a is not None else None,
                                          C.byref(rng), *tail))
    else:
        raise ValueError(name)
    return _motif_out(pw, npos, pos, n), st


def best_motif_information_content(variant: int, reps: int, src: Sources, m: int, k: int, pc: float,
                                   cutoff: float, rng, pcv=None, ppm=None, alphabet: bytes = DNA_BASES):
    n = src.n
    pw, npos, pos = _motif_state(n, None)
    n_out = C.c_int32()
    st = Stats()
    pcv_a = np.ascontiguousarray(pcv, np.float64) if pcv is not None else None
    ppm_a = np.ascontiguousarray(ppm, np.float64) if ppm is not None else None
    _check(lib().or_best_motif_information_content(
        C.c_int32(variant), C.c_int32(reps), _p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(n),
        C.c_int32(m), C.c_int32(k), C.c_double(pc), C.c_double(cutoff), _u8(alphabet), C.c_int32(len(alphabet)),
        _p(pcv_a, C.c_double) if pcv_a is not None else None,
        _p(ppm_a, C.c_double) if ppm_a is not None else None,
        C.byref(rng), _p(pw, C.c_double), _p(npos, C.c_int32), _p(pos, C.c_int32), C.byref(n_out), C.byref(st)))
    return _motif_out(pw, npos, pos, n_out.value), st


# ---------------------------------------------------------------------------------------------
# incremental mode (same arithmetic, no from-scratch rebuilds; see gibbs_oracle.c)
# ---------------------------------------------------------------------------------------------
PH_INIT, PH_GREEDY, PH_LEFT, PH_RIGHT = 1, 2, 4, 8


def fast_site_pipeline(variant: int, src: Sources, k: int, pc: float, *, pcv=None, rng=None, state=None,
                       phase_mask: int = 0, threads: int = 1, alphabet: bytes = DNA_BASES, log_cap: int = 4096):
    """variant 0 = WithBPV (fs:691), 1 = data-derived background (fs:697).
    Returns (scores, positions, Stats, sweep_log[(mode, moved, accepted)])."""
    n = src.n
    if state is not None:
        score = np.ascontiguousarray(state[0], np.float64).copy()
        pos = np.ascontiguousarray(state[1], np.int32).copy()
    else:
        score = np.zeros(n, np.float64)
        pos = np.zeros(n, np.int32)
    st = Stats()
    pcv_a = np.ascontiguousarray(pcv, np.float64) if pcv is not None else None
    log = np.zeros((log_cap, 3), np.int32)
    log_n = C.c_int32()
    _check(lib().or_fast_site_pipeline(
        C.c_int32(variant), C.c_int32(phase_mask), _p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(n),
        C.c_int32(k), C.c_double(pc), _u8(alphabet), C.c_int32(len(alphabet)),
        _p(pcv_a, C.c_double) if pcv_a is not None else None, C.byref(rng) if rng is not None else None,
        _p(score, C.c_double), _p(pos, C.c_int32), C.byref(st), C.c_int32(threads), _p(log, C.c_int32),
        C.c_int32(log_cap), C.byref(log_n)))
    return score, pos, st, [tuple(int(x) for x in r) for r in log[: log_n.value]]


def fast_site_chains(variant: int, src: Sources, k: int, pc: float, *, seed: int, chain_base: int, n_chains: int,
                     pcv=None, threads: int = 1, alphabet: bytes = DNA_BASES):
    """n_chains whole restarts with Philox streams (seed, chain_base + c). Returns (scores, positions, sums, Stats)."""
    n = src.n
    scores = np.zeros((n_chains, n), np.float64)
    pos = np.zeros((n_chains, n), np.int32)
    sums = np.zeros(n_chains, np.float64)
    st = Stats()
    pcv_a = np.ascontiguousarray(pcv, np.float64) if pcv is not None else None
    _check(lib().or_fast_site_chains(
        C.c_int32(variant), _p(src.buf, C.c_uint8), _p(src.off, C.c_int64), C.c_int32(n), C.c_int32(k),
        C.c_double(pc), _u8(alphabet), C.c_int32(len(alphabet)),
        _p(pcv_a, C.c_double) if pcv_a is not None else None, C.c_uint64(seed), C.c_int64(chain_base),
        C.c_int32(n_chains), C.c_int32(threads), _p(scores, C.c_double), _p(pos, C.c_int32), _p(sums, C.c_double),
        C.byref(st)))
    return scores, pos, sums, st
