"""Population-level entry points: one call flies every drone of a generation on the GPU.

This replaces the per-drone loop of the reference's evolve/evolution.py:64-73.  PyTorch is
used only to own device memory / streams; all arithmetic happens in libldrones.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._abi import GENOME_LEN, RECORD_DTYPE, STATE_LEN, LdParams, LdPidParams, make_params, make_pid_params

RECORD_WORDS = RECORD_DTYPE.itemsize // 8  # records travel as int64[n][4] on the torch side


def _params(params, kw) -> LdParams:
    if params is None:
        return make_params(**kw)
    if kw:
        raise TypeError("pass either `params` or fly() keyword arguments, not both")
    return params


# ---- host buffers in, host buffers out (C-ABI ld_fly_mlp_f64_host) ---------------------------
def _entry(L, precision: str, where: str):
    if precision not in ("f64", "f32"):
        raise ValueError("precision must be 'f64' (parity mode) or 'f32' (fast mode, statistical agreement only)")
    return getattr(L, f"ld_fly_mlp_{precision}_{where}")


def fly_population_host(genomes, params: LdParams = None, want_state: bool = False, out=None, precision: str = "f64", **kw):
    """genomes: (n, 314) float64 array-like in HOST memory -> (records[n] structured, state|None).

    `out` may be a preallocated RECORD_DTYPE array (e.g. backed by pinned memory).
    precision="f32" selects the float32 fast mode (no parity claim, see include/ldrones.h).
    """
    L = _lib.load()
    p = _params(params, kw)
    g = np.ascontiguousarray(genomes, dtype=np.float64)
    if g.ndim != 2 or g.shape[1] != GENOME_LEN:
        raise ValueError(f"genomes must have shape (n, {GENOME_LEN}), got {g.shape}")
    n = g.shape[0]
    rec = np.empty(n, dtype=RECORD_DTYPE) if out is None else out
    if rec.dtype != RECORD_DTYPE or rec.shape != (n,) or not rec.flags.c_contiguous:
        raise ValueError("`out` must be a contiguous RECORD_DTYPE array of length n")
    st = np.empty((n, STATE_LEN), dtype=np.float64) if want_state else None
    _lib.check(
        _entry(L, precision, "host")(C.byref(p), g.ctypes.data, n, rec.ctypes.data, st.ctypes.data if want_state else None)
    )
    return rec, st


# ---- device tensors (torch owns the memory) ---------------------------------------------------
def fly_population_device(genomes, params: LdParams = None, want_state: bool = False, records=None, state=None,
                          work_counter=None, precision: str = "f64", **kw):
    """genomes: CUDA float64 tensor (n, 314).  Asynchronous on torch's current stream.

    Returns (records int64[n][4], state float64[n][28] | None); decode with `records_to_numpy`
    or the `record_*` views below.
    """
    import torch

    L = _lib.load()
    p = _params(params, kw)
    if not (genomes.is_cuda and genomes.dtype == torch.float64 and genomes.dim() == 2 and genomes.shape[1] == GENOME_LEN):
        raise ValueError(f"genomes must be a CUDA float64 tensor of shape (n, {GENOME_LEN})")
    if not genomes.is_contiguous():
        raise ValueError("genomes must be contiguous (row = one drone's flattened brain)")
    n = genomes.shape[0]
    dev = genomes.device
    with torch.cuda.device(dev):
        if records is None:
            records = torch.empty((n, RECORD_WORDS), dtype=torch.int64, device=dev)
        if want_state and state is None:
            state = torch.empty((n, STATE_LEN), dtype=torch.float64, device=dev)
        if work_counter is None:
            work_counter = torch.empty(1, dtype=torch.int64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(
            _entry(L, precision, "device")(
                C.byref(p),
                genomes.data_ptr(),
                n,
                records.data_ptr(),
                state.data_ptr() if want_state else None,
                work_counter.data_ptr(),
                stream,
            )
        )
    return records, (state if want_state else None)


def records_to_numpy(records) -> np.ndarray:
    """int64[n][4] CUDA/CPU tensor -> structured numpy array (synchronises)."""
    a = records.detach().cpu().numpy()
    return a.view(RECORD_DTYPE).reshape(-1)


def record_reward(records):
    return records[:, 0]


def record_flight_time(records):
    import torch

    return records.view(torch.int32)[:, 3]


def record_coins(records):
    import torch

    return records.view(torch.int32)[:, 2]


def reduce_records_device(records, stats=None):
    """Warp-shuffle reduction on the device -> int64[5] CUDA tensor (active_steps, best_reward, best_index,
    coins, crashed); asynchronous on the current stream."""
    import torch

    L = _lib.load()
    n = records.shape[0]
    with torch.cuda.device(records.device):
        if stats is None:
            stats = torch.empty(5, dtype=torch.int64, device=records.device)
        stream = torch.cuda.current_stream(records.device).cuda_stream
        _lib.check(L.ld_reduce_records_device(records.data_ptr(), n, stats.data_ptr(), stream))
    return stats


def reduce_records(records):
    """Warp-shuffle reduction on the device -> dict(active_steps, best_reward, best_index, coins, crashed)."""
    s = reduce_records_device(records).cpu().tolist()
    return dict(active_steps=s[0], best_reward=s[1], best_index=s[2], coins=s[3], crashed=s[4])


def trace_drone(genome, params: LdParams = None, **kw):
    """Debug tap: fly one drone, return (record, trace[flight_time][28]) with the state after every step."""
    import torch

    L = _lib.load()
    p = _params(params, kw)
    g = torch.as_tensor(np.ascontiguousarray(genome, dtype=np.float64).reshape(GENOME_LEN)).cuda()
    rec = torch.zeros((1, RECORD_WORDS), dtype=torch.int64, device=g.device)
    tr = torch.full((max(int(p.timesteps), 1), STATE_LEN), float("nan"), dtype=torch.float64, device=g.device)
    with torch.cuda.device(g.device):
        stream = torch.cuda.current_stream(g.device).cuda_stream
        _lib.check(L.ld_fly_mlp_f64_trace_device(C.byref(p), g.data_ptr(), rec.data_ptr(), tr.data_ptr(), stream))
    r = records_to_numpy(rec)[0]
    return r, tr.cpu().numpy()[: int(r["flight_time"])]


# ---- PID controller mode (flight/collect_coins.py as a batch over gain sets) ---------------------
def fly_pid_population_host(pid_gains, params: LdParams = None, pid: LdPidParams = None, want_state: bool = False):
    """pid_gains: (n, 6, 3) per-controller (kp, ki, kd) — see _abi.paramset_to_pid_gains — in HOST memory.
    Returns (records[n], state|None); records['reward'] is the script's final score."""
    L = _lib.load()
    p = make_params() if params is None else params
    q = make_pid_params() if pid is None else pid
    g = np.ascontiguousarray(pid_gains, dtype=np.float64)
    if g.ndim != 3 or g.shape[1:] != (6, 3):
        raise ValueError(f"pid_gains must have shape (n, 6, 3), got {g.shape}")
    n = g.shape[0]
    rec = np.empty(n, dtype=RECORD_DTYPE)
    st = np.empty((n, STATE_LEN), dtype=np.float64) if want_state else None
    _lib.check(L.ld_fly_pid_f64_host(C.byref(p), C.byref(q), g.ctypes.data, n, rec.ctypes.data, st.ctypes.data if want_state else None))
    return rec, st


def fly_pid_population_device(pid_gains, params: LdParams = None, pid: LdPidParams = None, want_state: bool = False):
    """CUDA float64 tensor (n, 6, 3) -> (records int64[n][4], state|None); asynchronous on the current stream."""
    import torch

    L = _lib.load()
    p = make_params() if params is None else params
    q = make_pid_params() if pid is None else pid
    if not (pid_gains.is_cuda and pid_gains.dtype == torch.float64 and pid_gains.is_contiguous() and tuple(pid_gains.shape[1:]) == (6, 3)):
        raise ValueError("pid_gains must be a contiguous CUDA float64 tensor of shape (n, 6, 3)")
    n, dev = pid_gains.shape[0], pid_gains.device
    with torch.cuda.device(dev):
        records = torch.empty((n, RECORD_WORDS), dtype=torch.int64, device=dev)
        state = torch.empty((n, STATE_LEN), dtype=torch.float64, device=dev) if want_state else None
        _lib.check(L.ld_fly_pid_f64_device(C.byref(p), C.byref(q), pid_gains.data_ptr(), n, records.data_ptr(),
                                           state.data_ptr() if want_state else None, torch.cuda.current_stream(dev).cuda_stream))
    return records, state
