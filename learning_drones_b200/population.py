"""Population-level entry points: one call flies every drone of a generation on the GPU.

This replaces the per-drone loop of the reference's evolve/evolution.py:64-73.  PyTorch is
used only to own device memory / streams; all arithmetic happens in libldrones.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._abi import GENOME_LEN, RECORD_DTYPE, STATE_LEN, LdParams, LdPidParams, make_params, make_pid_params, paramset_to_pid_gains

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


def fly_population_host(genomes, params: LdParams = None, want_state: bool = False, out=None, precision: str = "f64",
                        init_state=None, **kw):
    """genomes: (n, 314) float64 array-like in HOST memory -> (records[n] structured, state|None).

    `out` may be a preallocated RECORD_DTYPE array (e.g. backed by pinned memory).
    precision="f32" selects the float32 fast mode (no parity claim, see include/ldrones.h).
    init_state: optional (n, 28) per-drone start rows (ld_fly_mlp_f64_from_host) replacing params.init_*.
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
    if init_state is not None:
        if precision != "f64":
            raise ValueError("per-drone start states are an fp64 (parity mode) feature")
        s0 = np.ascontiguousarray(init_state, dtype=np.float64)
        if s0.shape != (n, STATE_LEN):
            raise ValueError(f"init_state must have shape ({n}, {STATE_LEN}), got {s0.shape}")
        _lib.check(L.ld_fly_mlp_f64_from_host(C.byref(p), g.ctypes.data, s0.ctypes.data, n, rec.ctypes.data,
                                              st.ctypes.data if want_state else None))
        return rec, st
    _lib.check(
        _entry(L, precision, "host")(C.byref(p), g.ctypes.data, n, rec.ctypes.data, st.ctypes.data if want_state else None)
    )
    return rec, st


class HostSubmission:
    """A population submitted with `fly_population_host_submit`; `.result()` waits for it."""

    def __init__(self, ticket, genomes, records, state, init_state):
        self._ticket, self._done = ticket, False
        self._keep = (genomes, init_state)  # the library reads these until the wait returns
        self.records, self.state = records, state

    def result(self):
        """(records, state|None) once they are in host memory; raises LdError if the launch failed."""
        if not self._done:
            self._done = True
            _lib.check(_lib.load().ld_fly_mlp_f64_host_wait(self._ticket))
            self._keep = None
        return self.records, self.state


def fly_population_host_submit(genomes, params: LdParams = None, want_state: bool = False, out=None, init_state=None, **kw) -> HostSubmission:
    """Pipelined form of `fly_population_host` (ld_fly_mlp_f64_host_submit / _wait): returns at once; up to two
    submissions may be in flight per thread, and while one flies the next one's rows are already streaming to the GPU
    (a step then costs max(kernel, copy) instead of their sum).  For jobs that evaluate one population after another;
    a dependent generation loop keeps the population on the device instead (`evolve.gpu.GpuEvolution`).
    `genomes` / `out` should be pinned host memory for the copies to overlap."""
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
    s0 = None
    if init_state is not None:
        s0 = np.ascontiguousarray(init_state, dtype=np.float64)
        if s0.shape != (n, STATE_LEN):
            raise ValueError(f"init_state must have shape ({n}, {STATE_LEN}), got {s0.shape}")
    ticket = C.c_int64(0)
    _lib.check(L.ld_fly_mlp_f64_host_submit(C.byref(p), g.ctypes.data, s0.ctypes.data if s0 is not None else None, n, rec.ctypes.data,
                                            st.ctypes.data if want_state else None, C.byref(ticket)))
    return HostSubmission(ticket.value, g, rec, st, s0)


# ---- device tensors (torch owns the memory) ---------------------------------------------------
def fly_population_device(genomes, params: LdParams = None, want_state: bool = False, records=None, state=None,
                          work_counter=None, precision: str = "f64", init_state=None, **kw):
    """genomes: CUDA float64 tensor (n, 314).  Asynchronous on torch's current stream.

    Returns (records int64[n][4], state float64[n][28] | None); decode with `records_to_numpy`
    or the `record_*` views below.  `work_counter`: int64[2] CUDA scratch (queue head, error flag); pass your
    own to look at the flag with `check_work_counter` once the stream has been synchronised.
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
            work_counter = torch.empty(2, dtype=torch.int64, device=dev)
        if work_counter.numel() < 2 or work_counter.dtype != torch.int64 or not work_counter.is_cuda:
            raise ValueError("work_counter must be a CUDA int64 tensor with 2 elements (queue head, error flag)")
        stream = torch.cuda.current_stream(dev).cuda_stream
        if init_state is not None:
            if precision != "f64":
                raise ValueError("per-drone start states are an fp64 (parity mode) feature")
            if not (init_state.is_cuda and init_state.dtype == torch.float64 and init_state.is_contiguous()
                    and tuple(init_state.shape) == (n, STATE_LEN)):
                raise ValueError(f"init_state must be a contiguous CUDA float64 tensor of shape ({n}, {STATE_LEN})")
            _lib.check(L.ld_fly_mlp_f64_from_device(C.byref(p), genomes.data_ptr(), init_state.data_ptr(), n, records.data_ptr(),
                                                    state.data_ptr() if want_state else None, work_counter.data_ptr(), stream))
        else:
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


def check_work_counter(work_counter):
    """After the launch has completed: raise if the episode kernel gave up waiting for its work queue
    (work_counter[1], see include/ldrones.h).  Synchronises."""
    word = int(work_counter[1].item()) & 0xFFFFFFFF
    if word != 0:
        raise _lib.LdError(-2, f"episode kernel error word {word} (1 = gave up waiting for its work queue; >= 2 = a bounds check of the "
                               "checked build failed); records are incomplete")


PROFILE_SLOTS = ("claim", "stage", "wait", "hot", "post", "stage_events", "lanes_staged", "hot_entries", "warp_steps",
                 "live_lane_steps", "warps", "stage_state", "stage_copy", "stage_tmem", "stage_groups", "slow_trig_steps")


def fly_population_profile(genomes, params: LdParams = None, timeline: bool = False, **kw):
    """Development tap (ld_debug_fly_profile_device): fly `genomes` (CUDA float64 (n, 314)) with the instrumented
    kernel and return (records, dict) — cycles per phase summed over all warps and event counts; with timeline=True
    the dict also holds "times": int64 (n, 2) ns (first start, landing) relative to the earliest start.  Synchronises."""
    import torch

    L = _lib.load()
    p = _params(params, kw)
    n, dev = genomes.shape[0], genomes.device
    with torch.cuda.device(dev):
        records = torch.empty((n, RECORD_WORDS), dtype=torch.int64, device=dev)
        counter = torch.empty(2, dtype=torch.int64, device=dev)
        prof = torch.zeros(16, dtype=torch.int64, device=dev)
        times = torch.zeros((n, 2), dtype=torch.int64, device=dev) if timeline else None
        _lib.check(L.ld_debug_fly_profile_device(C.byref(p), genomes.data_ptr(), n, records.data_ptr(), counter.data_ptr(), prof.data_ptr(),
                                                 times.data_ptr() if timeline else None, torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize(dev)
    out = dict(zip(PROFILE_SLOTS, prof.cpu().tolist()))
    if timeline:
        t = times.cpu().numpy()
        out["times"] = t - t[:, 0].min()
    return records, out


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
def paramset_gains(paramset) -> np.ndarray:
    """Gain set(s) of the reference's table by name — "11" or ["00", "11", ...] (drone/quadcopter/choose_paramset.py,
    `run_simulation_without_viz.sh 11`) -> (n, 6, 3) per-controller (kp, ki, kd) wired like collect_coins.py:125-188."""
    from .drone.quadcopter.choose_paramset import convert_paramset_2_float, get_paramset

    names = [paramset] if isinstance(paramset, (str, int)) else list(paramset)
    return paramset_to_pid_gains(np.array([convert_paramset_2_float(get_paramset(str(n))) for n in names], dtype=np.float64))


def fly_pid_population_host(pid_gains=None, params: LdParams = None, pid: LdPidParams = None, want_state: bool = False, paramset=None):
    """pid_gains: (n, 6, 3) per-controller (kp, ki, kd) — see _abi.paramset_to_pid_gains — in HOST memory, or
    `paramset="11"` / a list of names from the reference's gain table (config 1 = paramset "11").
    Returns (records[n], state|None); records['reward'] is the script's final score."""
    L = _lib.load()
    p = make_params() if params is None else params
    q = make_pid_params() if pid is None else pid
    if (pid_gains is None) == (paramset is None):
        raise TypeError("pass either pid_gains or paramset")
    if paramset is not None:
        pid_gains = paramset_gains(paramset)
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
