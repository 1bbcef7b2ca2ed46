"""ctypes wrapper around oracle/libld_oracle.so — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
import this module.  It reuses the product's ctypes struct definitions
(learning_drones_b200._abi) so that both sides agree on the ld_params layout; the
arithmetic lives in ld_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from learning_drones_b200._abi import GENOME_LEN, RECORD_DTYPE, STATE_LEN, LdGaParams, LdParams, LdPidParams, LdRecord

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libld_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ld_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ldrones.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(_LIB_PATH) for f in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libld_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        L.ld_oracle_fly_mlp.argtypes = [C.POINTER(LdParams), dp, C.c_int64, C.POINTER(LdRecord), dp, C.c_int, C.c_int]
        L.ld_oracle_fly_mlp.restype = C.c_int
        L.ld_oracle_fly_mlp_from.argtypes = [C.POINTER(LdParams), dp, dp, C.c_int64, C.POINTER(LdRecord), dp, C.c_int, C.c_int]
        L.ld_oracle_fly_mlp_from.restype = C.c_int
        L.ld_oracle_trace_mlp.argtypes = [C.POINTER(LdParams), dp, C.POINTER(LdRecord), dp, C.c_int]
        L.ld_oracle_trace_mlp.restype = C.c_int
        L.ld_oracle_accelerations.argtypes = [C.POINTER(LdParams), dp, dp, dp, dp, C.c_int]
        L.ld_oracle_accelerations.restype = C.c_int
        L.ld_oracle_max_threads.restype = C.c_int
        L.ld_oracle_fly_pid.argtypes = [C.POINTER(LdParams), C.POINTER(LdPidParams), dp, C.c_int64, C.POINTER(LdRecord), dp, dp, C.c_int, C.c_int]
        L.ld_oracle_fly_pid.restype = C.c_int
        L.ld_oracle_ga_step.argtypes = [C.POINTER(LdGaParams), dp, C.POINTER(C.c_int64), C.c_int64, dp, C.c_int64, C.c_int64,
                                        C.POINTER(C.c_int32)]
        L.ld_oracle_ga_step.restype = C.c_int
        _lib = L
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def max_threads() -> int:
    return int(lib().ld_oracle_max_threads())


def fly_mlp(params: LdParams, genomes, use_fma: bool = True, nthreads: int = 0, want_state: bool = False, init_state=None):
    """Population loop around the reference's fly() — returns (records, final_state|None).
    init_state: optional (n, 28) rows the Drones hold when fly() is entered (counters of the episode start at 0)."""
    g = np.ascontiguousarray(genomes, dtype=np.float64).reshape(-1, GENOME_LEN)
    n = g.shape[0]
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    st = np.zeros((n, STATE_LEN), dtype=np.float64) if want_state else None
    s0 = None if init_state is None else np.ascontiguousarray(init_state, dtype=np.float64).reshape(n, STATE_LEN)
    rc = lib().ld_oracle_fly_mlp_from(
        C.byref(params),
        _dptr(g),
        _dptr(s0) if s0 is not None else None,
        n,
        rec.ctypes.data_as(C.POINTER(LdRecord)),
        _dptr(st) if want_state else None,
        int(use_fma),
        int(nthreads),
    )
    if rc != 0:
        raise RuntimeError(f"ld_oracle_fly_mlp failed: {rc}")
    return rec, st


def trace_mlp(params: LdParams, genome, use_fma: bool = True):
    """One drone; returns (record, trace[timesteps][STATE_LEN]) — rows past the last step are NaN."""
    g = np.ascontiguousarray(genome, dtype=np.float64).reshape(GENOME_LEN)
    rec = np.zeros(1, dtype=RECORD_DTYPE)
    tr = np.full((max(int(params.timesteps), 1), STATE_LEN), np.nan)
    rc = lib().ld_oracle_trace_mlp(C.byref(params), _dptr(g), rec.ctypes.data_as(C.POINTER(LdRecord)), _dptr(tr), int(use_fma))
    if rc != 0:
        raise RuntimeError(f"ld_oracle_trace_mlp failed: {rc}")
    return rec[0], tr


def accelerations(params: LdParams, angle, thrust, use_fma: bool = True):
    a = np.ascontiguousarray(angle, dtype=np.float64).reshape(3)
    t = np.ascontiguousarray(thrust, dtype=np.float64).reshape(4)
    lin = np.zeros(3)
    rot = np.zeros(3)
    lib().ld_oracle_accelerations(C.byref(params), _dptr(a), _dptr(t), _dptr(lin), _dptr(rot), int(use_fma))
    return lin, rot


def fly_pid(params: LdParams, pid: LdPidParams, pid_gains, use_fma: bool = True, nthreads: int = 0, want_state: bool = False,
            want_trace: bool = False):
    """Batch of flight/collect_coins.py episodes; pid_gains (n, 6, 3). Returns (records, state|None, trace|None)."""
    g = np.ascontiguousarray(pid_gains, dtype=np.float64).reshape(-1, 18)
    n = g.shape[0]
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    st = np.zeros((n, STATE_LEN)) if want_state else None
    tr = np.full((int(pid.timesteps), STATE_LEN), np.nan) if want_trace else None
    rc = lib().ld_oracle_fly_pid(C.byref(params), C.byref(pid), _dptr(g), n, rec.ctypes.data_as(C.POINTER(LdRecord)),
                                 _dptr(st) if want_state else None, _dptr(tr) if want_trace else None, int(use_fma), int(nthreads))
    if rc != 0:
        raise RuntimeError(f"ld_oracle_fly_pid failed: {rc}")
    return rec, st, tr


def ga_step(ga: LdGaParams, parents, rewards, lo: int = 0, hi: int = None):
    """evolve/util.selection_and_crossover + Drone.mutate on the genome matrix with the library's
    Philox stream. Returns (children[hi-lo][314], picks[hi-lo][5] = parent, cut0..2 (-1 none), n_mutated)."""
    g = np.ascontiguousarray(parents, dtype=np.float64).reshape(-1, GENOME_LEN)
    n = g.shape[0]
    hi = n if hi is None else hi
    r = np.ascontiguousarray(rewards, dtype=np.int64)
    out = np.zeros((hi - lo, GENOME_LEN))
    picks = np.zeros((hi - lo, 5), dtype=np.int32)
    rc = lib().ld_oracle_ga_step(C.byref(ga), _dptr(g), r.ctypes.data_as(C.POINTER(C.c_int64)), n, _dptr(out), lo, hi,
                                 picks.ctypes.data_as(C.POINTER(C.c_int32)))
    if rc != 0:
        raise RuntimeError(f"ld_oracle_ga_step failed: {rc}")
    return out, picks
