#!/usr/bin/env python
"""bench.py — drone-steps/s of the population fitness evaluation (fly() path) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--population survivors|random] [--drones-per-gpu D] [--timesteps H]

A "step" is one pass of the hot path over one batch: every drone of the per-GPU population flies one H-timestep
episode (one generation's fitness evaluation).  Workload at N=1 = BASELINE.json configs[2]: 65,536 synthetic
drones, 2,000-step coin episode, the all-survivor (S) family of SURVEY.md §8d so that active == nominal
drone-steps.  N>1 (torchrun, one rank per GPU): each rank flies its own 65,536-drone shard (weak scaling, drones
are independent) and the fitness records are all-gathered with NCCL inside the timed step; the same JSON line
then also carries BASELINE.json configs[3] (`config4`: 1,048,576 drones split N ways) and configs[4] (`config5`:
100 generations of the full loop on that population).

Prints ONE JSON line (rank 0).  `--impl reference` times the REFERENCE'S OWN fly() (unmodified Python files
staged under oracle/_ref by oracle/build_ref.sh) under multiprocessing.Pool on all host cores, on a bounded
sample of the same workload; if those files are absent it falls back to the C restatement (oracle port).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "drone-steps/sec (dynamics+brain+reward)"
UNIT = "drone-steps/s"
FLOP_PER_STEP = 850.0  # SURVEY.md §8d algorithmic work per drone-step (FMA = 2)
GENOME_BYTES = 314 * 8
RECORD_BYTES = 32
# bounded CPU samples (fixed, so that both arms print the same `config`): drones of the workload population
REF_ARM_DRONES_PER_STEP = 128   # `--impl reference`: per step, reference fly() under Pool(all cores)
CPU_BASELINE_DRONES = 768       # our arm's one-shot `cpu_baseline` (reference fly(): ~20 s on 16 cores)
PORT_BASELINE_DRONES = 32768    # C restatement (oracle port): ~1 s per 16 cores
CONFIG4_DRONES = 1 << 20


def population(kind, n, seed=2026, lo=0, hi=None):
    """The config-3 generators of SURVEY.md §8d (same formulas as oracle/gen_golden.py; the golden fixtures are
    the first rows of exactly these arrays for seed 2026).  Rows [lo, hi) of the n-drone population."""
    hi = n if hi is None else hi
    rng = np.random.default_rng(seed)
    hover_bias = float(np.log(0.1962 / 0.8038) / 0.4)
    # draw both families in the fixture's order so that prefixes match tests/golden
    delta = rng.uniform(-0.010, 0.100, size=n)[lo:hi]
    if kind == "survivors":
        g_s = np.zeros((hi - lo, 314))
        for j in range(4):
            g_s[:, 270 + j * 11 + 10] = hover_bias + delta
        return g_s
    cols = [rng.normal(0.0, 1.0 / np.sqrt(r), size=(n, r * c))[lo:hi] for r, c in ((10, 16), (10, 11), (4, 11))]
    return np.concatenate(cols, axis=1)


def random_shard(n_local, rank, seed=2026):
    """Reference init N(0, 1/sqrt(n_out)) per layer (evolve/brain.py:24) for ONE rank's shard of a big population
    (config 5); every rank draws its own stream so that no rank materialises the whole 2.6 GB matrix."""
    rng = np.random.default_rng([seed, rank])
    return np.concatenate([rng.normal(0.0, 1.0 / np.sqrt(r), size=(n_local, r * c)) for r, c in ((10, 16), (10, 11), (4, 11))], axis=1)


class ClockSampler:
    """SM clock / throttle-reason sampler running DURING the timed region (NVML every ~5 ms; if NVML delivers
    nothing, `nvidia-smi -lms 20`)."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.threads, self.max_mhz, self.source = index, [], False, [], None, None
        self.proc = None

    def _loop_nvml(self):
        try:
            import pynvml as N

            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        except Exception:  # noqa: BLE001 - any NVML problem -> the nvidia-smi loop takes over
            return
        while not self.stop_flag:
            try:
                self.rows.append((time.perf_counter(), float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)),
                                  N.nvmlDeviceGetPowerUsage(h) / 1000.0, int(get_reasons(h))))
                self.source = self.source or "nvml"
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def _loop_smi(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        for line in self.proc.stdout:
            if self.stop_flag:
                break
            f = [x.strip() for x in line.split(",")]
            try:
                row = (time.perf_counter(), float(f[0]), float(f[2]), int(f[3], 16))
            except (ValueError, IndexError):
                continue
            if self.source in (None, "nvidia-smi"):  # NVML rows win when both deliver
                self.max_mhz = float(f[1])
                self.rows.append(row)
                self.source = "nvidia-smi"

    def start(self, wait_s=3.0):
        for target in (self._loop_nvml, self._loop_smi):
            t = threading.Thread(target=target, daemon=True)
            t.start()
            self.threads.append(t)
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < wait_s:  # the timed region starts with the sampler delivering
            time.sleep(0.01)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        for t in self.threads:
            t.join(timeout=1.0)
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        rows = inside or self.rows[-3:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "no sampler delivered (NVML and nvidia-smi both failed)"}
        sm = sorted(r[1] for r in rows)
        mask = 0
        for r in rows:
            mask |= r[3]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "power_w_max": max(r[2] for r in rows),
                "samples": len(rows), "samples_inside_timed_region": len(inside), "source": self.source,
                "reasons": [name for bit, name in self.REASONS if mask & bit]}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def csrc_hash():
    """sha256 over the kernel sources: numbers taken from an ncu capture are only valid for the code they were captured on."""
    h = hashlib.sha256()
    base = os.path.join(ROOT, "learning_drones_b200", "csrc")
    for name in sorted(os.listdir(base)) + ["../../include/ldrones.h"]:
        with open(os.path.join(base, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def kernel_stamp(D, H, pop):
    """profiles/kernel_stamp.json: DRAM traffic and executed fp64-pipe instructions per drone-step of
    fly_mlp_f64_kernel from ncu captures, stamped with the capture files, the workload and the csrc hash."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_stamp.json")) as f:
            st = json.load(f)
    except (OSError, ValueError):
        return None, None, "profiles/kernel_stamp.json missing"
    if st.get("csrc_sha16") != csrc_hash():
        return None, None, f"stale: profiles/kernel_stamp.json was captured on csrc {st.get('csrc_sha16')}, this build is {csrc_hash()}"
    w = st.get("workload", {})
    same = (w.get("drones"), w.get("timesteps"), w.get("population")) == (D, H, pop)
    return (st.get("traffic_bytes") if same else None), st.get("fp64_instr_per_drone_step"), st.get("source")


# ---- CPU baselines -------------------------------------------------------------------------------------------------
def port_baseline(kind, H, sample, n_total, threads=0):
    """The C restatement of the reference (oracle port) on the host cores, bounded sample."""
    from learning_drones_b200._abi import make_params
    from oracle import oracle as O

    g = population(kind, max(n_total, sample), hi=sample)
    p = make_params(H)
    O.fly_mlp(p, g[:64], nthreads=threads)  # warm
    t0 = time.perf_counter()
    rec, _ = O.fly_mlp(p, g, nthreads=threads)
    dt = time.perf_counter() - t0
    steps = int(rec["flight_time"].sum())
    cores = threads or O.max_threads()
    return {"value": steps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample} drones of the workload population x {H} steps ({steps} active drone-steps, {dt:.2f} s); "
                      f"C restatement oracle/ld_oracle.c on {cores} threads"}, rec


def reference_available():
    from oracle import ref_runner

    return ref_runner.reference_root() is not None


def reference_baseline(pool, kind, H, sample, n_total):
    """The reference's own fly() under multiprocessing.Pool, bounded sample."""
    g = population(kind, max(n_total, sample), hi=sample)
    r = pool.fly(g, H)
    return {"value": r["active_steps"] / r["seconds"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
            "per_core": r["active_steps"] / r["seconds"] / r["cores"],
            "sample": f"first {sample} drones of the workload population x {H} steps ({r['active_steps']} active drone-steps, "
                      f"{r['seconds']:.1f} s); the reference's unmodified flight.fly() (oracle/_ref), one drone per task, "
                      f"multiprocessing.Pool({r['cores']})"}, r


def workload_config(args, world):
    """`config` of the JSON line — a-priori facts only, identical for our arm and the reference arm."""
    D, H = args.drones_per_gpu, args.timesteps
    return {"workload": f"configs[2]: {D}-drone {args.population} population per GPU, {H}-step coin episode, "
                        "batched fitness eval" + (" + NCCL all-gather of fitness records" if world > 1 else ""),
            "population": args.population, "timesteps": H, "drones_per_gpu": D, "global_drones": D * world,
            "nominal_drone_steps_per_step": D * world * H,
            "l2": f"inputs larger than L2 ({D * GENOME_BYTES / 1e6:.0f} MB genome matrix per GPU re-read every step)",
            "parallelism": f"population sharded over {world} GPU(s), no data-path collective; records all-gathered",
            "schedule": ("ld_params.slice_steps = LD_SLICE_STAGGER (scheduling only: staggered first wave, for populations of full-horizon "
                         "drones; host-buffer launches use the library default)" if args.population == "survivors" else "library default"),
            "cpu_samples": {"reference_arm_drones_per_step": REF_ARM_DRONES_PER_STEP, "cpu_baseline_drones": CPU_BASELINE_DRONES,
                            "port_baseline_drones": PORT_BASELINE_DRONES,
                            "note": "the CPU arms fly the first rows of the same population; rates are per active drone-step"}}


def run_reference(args, rank, world):
    """`--impl reference`: rank 0 alone works; the other ranks exit 0."""
    if rank != 0:
        return
    H, n_total = args.timesteps, args.drones_per_gpu * world
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world), "gpu_launches": 0}
    if reference_available() and not args.port_only:
        from oracle.ref_runner import ReferencePool

        pool = ReferencePool()
        sample = args.cpu_sample or REF_ARM_DRONES_PER_STEP
        for _ in range(args.warmup):
            pool.fly(population(args.population, n_total, hi=min(sample, 2 * pool.nproc)), H)
        steps, secs, cores = 0, 0.0, pool.nproc
        for _ in range(args.steps):
            base, r = reference_baseline(pool, args.population, H, sample, n_total)
            steps += r["active_steps"]
            secs += r["seconds"]
        pool.close()
        value = steps / secs
        line["cpu_baseline"] = {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "per_core": value / cores,
                                "sample": f"each step: first {sample} drones of the workload population x {H} steps through the reference's "
                                          f"unmodified flight.fly() (oracle/_ref), multiprocessing.Pool({cores}); {steps} active drone-steps in {secs:.1f} s"}
        port, _ = port_baseline(args.population, H, PORT_BASELINE_DRONES, n_total)
        line["cpu_baseline_port"] = port
    else:
        sample = args.cpu_sample or PORT_BASELINE_DRONES
        for _ in range(args.warmup):
            port_baseline(args.population, H, 512, n_total)
        t0, steps = time.perf_counter(), 0
        for _ in range(args.steps):
            port, rec = port_baseline(args.population, H, sample, n_total)
            steps += int(rec["flight_time"].sum())
        secs = time.perf_counter() - t0
        value = steps / secs
        port["value"] = value
        port["sample"] = f"each step: {port['sample']}; oracle/_ref (the reference's Python files) is not staged on this box"
        line["cpu_baseline"] = port
    line["value"] = value
    line["ms_per_step"] = 1e3 * secs / args.steps
    line["e2e"] = {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    print(json.dumps(line))


# ---- topology: put each rank on the cores next to its GPU ----------------------------------------------------------------
def bind_to_gpu_cores(local_rank, local_world):
    """Best effort: `nvidia-smi topo -m` lists the CPU affinity of every GPU; ranks whose GPUs share a core range
    split it.  Pinned staging buffers are then first-touched (and the copy-issuing thread runs) next to the GPU."""
    info = {"bound": False}
    try:
        out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
        rows = {}
        header = None
        for ln in out.splitlines():
            clean = ln.replace("\x1b[4m", "").replace("\x1b[0m", "")
            if "CPU Affinity" in clean and header is None:
                header = clean
                continue
            if clean.startswith("GPU") and header is not None:
                f = clean.split()
                idx = int(f[0][3:])
                rng = next((x for x in f[1:] if x[0].isdigit() and ("-" in x or "," in x or x.isdigit()) and not x.startswith("NV")), None)
                numa = next((x for x in reversed(f) if x.isdigit()), None)
                if rng:
                    rows[idx] = (rng, numa)
        if local_rank not in rows:
            return info
        rng, numa = rows[local_rank]
        cores = []
        for part in rng.split(","):
            a, _, b = part.partition("-")
            cores += list(range(int(a), int(b or a) + 1))
        sharing = sorted(i for i, (r, _) in rows.items() if r == rng and i < local_world)
        k = sharing.index(local_rank) if local_rank in sharing else 0
        per = max(len(cores) // max(len(sharing), 1), 1)
        mine = cores[k * per:(k + 1) * per] or cores
        allowed = sorted(set(mine) & set(os.sched_getaffinity(0))) or sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info = {"bound": True, "gpu_cpu_affinity": rng, "numa": numa, "cores": f"{allowed[0]}-{allowed[-1]}", "ranks_sharing_range": len(sharing)}
    except Exception as e:  # noqa: BLE001 - topology is advisory
        info["error"] = repr(e)
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--population", default="survivors", choices=["survivors", "random"])
    ap.add_argument("--drones-per-gpu", type=int, default=65536)
    ap.add_argument("--timesteps", type=int, default=2000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="drones in the CPU sample (default: the fixed sizes named in config.cpu_samples)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--port-only", action="store_true", help="CPU arms use the C restatement even if oracle/_ref is staged")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f64 = parity mode (the benchmark); f32 = fast mode, statistical agreement only, never the headline")
    ap.add_argument("--generations", type=int, default=5, help="generations of the full GA loop to time (0 = skip)")
    ap.add_argument("--no-config45", action="store_true", help="skip the config4 / config5 blocks of a multi-GPU run")
    ap.add_argument("--config5-generations", type=int, default=100)
    ap.add_argument("--no-bind", action="store_true", help="do not bind ranks to the cores next to their GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    binding = {"bound": False} if (args.no_bind or world == 1) else bind_to_gpu_cores(local_rank, local_world)

    import torch
    import torch.distributed as dist

    from learning_drones_b200 import _lib, population as ldpop
    from learning_drones_b200._abi import RECORD_DTYPE, make_params
    from learning_drones_b200.distributed import allgather_records

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return int(t.item())

    D, H = args.drones_per_gpu, args.timesteps
    prec = args.precision
    if prec != "f64":
        args.generations = 0

    def measure(host_genomes, params, steps, warmup, with_sampler):
        """Device-resident `value` and host-buffer `e2e` of one workload: (total_ms, kernel_ms, active, e2e_s, clocks)."""
        Dl = host_genomes.shape[0]
        pinned = torch.from_numpy(np.ascontiguousarray(host_genomes)).pin_memory()
        genomes = pinned.to(dev)
        records = torch.empty((Dl, 4), dtype=torch.int64, device=dev)
        counter = torch.empty(2, dtype=torch.int64, device=dev)  # queue head, error flag

        def step():
            ldpop.fly_population_device(genomes, params, records=records, work_counter=counter, precision=prec)
            if world > 1:
                return allgather_records(records, n_global=Dl * world)  # fitness all-gather for selection (NCCL/NVLink)
            return records

        for _ in range(warmup):
            step()
        barrier()
        sampler = ClockSampler(local_rank) if (with_sampler and rank == 0) else None
        if sampler:
            sampler.start()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        e0.record()
        for a, b in ev:
            a.record()
            ldpop.fly_population_device(genomes, params, records=records, work_counter=counter, precision=prec)
            b.record()
            if world > 1:
                allgather_records(records, n_global=Dl * world)
        e1.record()
        barrier()
        t_wall1 = time.perf_counter()
        ldpop.check_work_counter(counter)
        total_ms = max_over_ranks(e0.elapsed_time(e1))
        kernel_ms = sum(a.elapsed_time(b) for a, b in ev) / steps
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        rec_np = ldpop.records_to_numpy(records)
        active = sum_over_ranks(int(rec_np["flight_time"].sum()))  # per step, whole job
        # end to end through the host C-ABI call (pinned host buffers, copies inside)
        out_pinned = torch.empty((Dl, 4), dtype=torch.int64).pin_memory()
        out_np = out_pinned.numpy().view(RECORD_DTYPE).reshape(-1)
        host_np = pinned.numpy()
        ldpop.fly_population_host(host_np, params, out=out_np, precision=prec)  # warm (allocates the workspace)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            ldpop.fly_population_host(host_np, params, out=out_np, precision=prec)
            _ = int(out_np["reward"].max())  # the step's result is read on the host
        e2e_rank_s = time.perf_counter() - t0
        e2e_s = max_over_ranks(e2e_rank_s)
        assert out_np.tobytes() == rec_np.tobytes(), "host-API result differs from device-API result"
        # the same through the pipelined host API (depth 2: step k+1's rows stream in while step k flies); every step's
        # host->device copy and result read are inside the timed region here too
        pipe_s = None
        if prec == "f64":
            outs = [torch.empty((Dl, 4), dtype=torch.int64).pin_memory().numpy().view(RECORD_DTYPE).reshape(-1) for _ in range(2)]
            warm = [ldpop.fly_population_host_submit(host_np, params, out=outs[i]) for i in range(2)]  # warm: allocates BOTH workspaces
            for w_ in warm:
                w_.result()
            pipe_runs = []
            for _ in range(3):  # three runs of K steps each; the best one is reported, all are listed
                barrier()
                t0 = time.perf_counter()
                pending = None
                for k in range(steps):
                    nxt = ldpop.fly_population_host_submit(host_np, params, out=outs[k & 1])
                    if pending is not None:
                        _ = int(pending.result()[0]["reward"].max())
                    pending = nxt
                _ = int(pending.result()[0]["reward"].max())
                pipe_runs.append(max_over_ranks(time.perf_counter() - t0))
            pipe_s = min(pipe_runs)
            assert outs[(steps - 1) & 1].tobytes() == rec_np.tobytes(), "pipelined host-API result differs"
        # the host->device leg on its own, every rank at the same time (names the end-to-end limiter at N > 1)
        barrier()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        for _ in range(3):
            genomes.copy_(pinned, non_blocking=True)
        h1.record()
        torch.cuda.synchronize()
        h2d_gbs = 3 * Dl * GENOME_BYTES / (h0.elapsed_time(h1) * 1e-3) / 1e9
        return dict(total_ms=total_ms, kernel_ms=kernel_ms, active=active, e2e_s=e2e_s, e2e_rank_s=e2e_rank_s, clocks=clocks,
                    h2d_gbs=h2d_gbs, drones_local=Dl, pipe_s=pipe_s, pipe_runs=(pipe_runs if prec == "f64" else None))

    def gather_floats(x):
        if world == 1:
            return [x]
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    # ---- the headline workload --------------------------------------------------------------------------------------
    params = make_params(H)
    params.refill_batch = int(os.environ.get("LD_REFILL_BATCH", "0"))
    # scheduling only (never changes results): the (S) family flies every episode to the horizon, which is what the
    # staggered-first-wave schedule is for; mixed / random populations keep the library default
    from learning_drones_b200._abi import SLICE_STAGGER

    params.slice_steps = int(os.environ.get("LD_SLICE_STEPS", str(SLICE_STAGGER if args.population == "survivors" else 0)))
    m = measure(population(args.population, D * world, lo=rank * D, hi=(rank + 1) * D), params, args.steps, args.warmup, True)
    value = m["active"] * args.steps / (m["total_ms"] * 1e-3)
    e2e_value = m["active"] * args.steps / m["e2e_s"]
    h2d_all = gather_floats(m["h2d_gbs"])
    e2e_all = gather_floats(m["e2e_rank_s"])

    # ---- generations/s: full loop (fly + record all-gather + best tracking + GA step + elitism) --------------------
    def time_generations(evo, n_gen):
        evo.step()  # warm (allocations, NCCL)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        hist = evo.run(n_gen)
        g1.record()
        barrier()
        return max_over_ranks(g0.elapsed_time(g1)), hist

    gens = None
    if args.generations > 0:
        from learning_drones_b200.evolve.gpu import GpuEvolution

        evo = GpuEvolution(population(args.population, D * world, lo=rank * D, hi=(rank + 1) * D), seed=2026, timesteps=H,
                           local_shard=True, n_global=D * world)
        gms, hist = time_generations(evo, args.generations)
        gens = {"value": args.generations / (gms * 1e-3), "unit": "generations/s", "generations": args.generations,
                "population": D * world, "timesteps": H, "ga_mode": evo.ga_mode, "p2p_error": evo.p2p_error,
                "active_drone_steps_per_generation": [h["active_steps"] for h in hist],
                "best_score": [h["global_best_score"] for h in hist],
                "note": "one untimed warm-up generation flies the bench population; the timed generations are its GA offspring "
                        "(most crash early), enqueued back to back with one host synchronisation at the end"}
        del evo

    # ---- BASELINE configs[3] and [4] on the multi-GPU box ---------------------------------------------------------------
    config4 = config5 = None
    if world > 1 and prec == "f64" and not args.no_config45 and CONFIG4_DRONES % (2 * world) == 0:
        from learning_drones_b200.evolve.gpu import GpuEvolution

        D4 = CONFIG4_DRONES // world
        p4 = make_params(H)
        p4.slice_steps = SLICE_STAGGER
        m4 = measure(population("survivors", CONFIG4_DRONES, lo=rank * D4, hi=(rank + 1) * D4), p4, 5, 3, False)
        config4 = {"workload": f"configs[3]: {CONFIG4_DRONES}-drone survivors population sharded over {world} GPUs ({D4} per GPU), "
                               f"{H}-step episode + NCCL all-gather of the fitness records",
                   "value": m4["active"] * 5 / (m4["total_ms"] * 1e-3), "unit": UNIT, "ms_per_step": m4["total_ms"] / 5, "steps": 5, "warmup": 3,
                   "kernel_ms": m4["kernel_ms"], "active_drone_steps_per_step": m4["active"],
                   "e2e": {"value": m4["active"] * 5 / m4["e2e_s"], "unit": UNIT, "h2d_bytes_per_step": CONFIG4_DRONES * GENOME_BYTES,
                           "d2h_bytes_per_step": CONFIG4_DRONES * RECORD_BYTES,
                           "pipelined_value": None if m4["pipe_s"] is None else m4["active"] * 5 / m4["pipe_s"]},
                   "north_star_target": 1e10}
        evo = GpuEvolution(random_shard(D4, rank), seed=2026, timesteps=H, local_shard=True, n_global=CONFIG4_DRONES)
        g5 = args.config5_generations
        gms, hist = time_generations(evo, g5)
        config5 = {"workload": f"configs[4]: full neuroevolution loop (fly + record all-gather + best tracking + selection/crossover/mutation "
                               f"+ elitism) for {g5} generations, {CONFIG4_DRONES} drones (reference init N(0, 1/sqrt(n_out))) on {world} GPUs",
                   "seconds": gms * 1e-3, "generations": g5, "value": g5 / (gms * 1e-3), "unit": "generations/s",
                   "ga_mode": evo.ga_mode, "p2p_error": evo.p2p_error,
                   "active_drone_steps": int(sum(h["active_steps"] for h in hist)),
                   "drone_steps_per_s": sum(h["active_steps"] for h in hist) / (gms * 1e-3),
                   "best_score_first_last": [hist[0]["global_best_score"], hist[-1]["global_best_score"]],
                   "note": "after one untimed warm-up generation; generations enqueued back to back, one host synchronisation at the end"}
        del evo

    if rank == 0:
        import ctypes

        peaks, peaks_src = measured_peaks()
        tf, per_clk = ctypes.c_double(), ctypes.c_double()
        _lib.check(L.ld_measure_fp64_peak(tf, per_clk))
        kernel_ms = m["kernel_ms"]
        alg_bytes = D * (GENOME_BYTES + RECORD_BYTES)  # per launch: genome rows in, records out
        achieved_gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
        steps_per_launch = m["active"] / world
        fp64_tf = steps_per_launch * FLOP_PER_STEP / (kernel_ms * 1e-3) / 1e12
        traffic, instr, stamp_src = kernel_stamp(D, H, args.population)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": m["total_ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec, "data": "synthetic",
            "config": workload_config(args, world),
            "active_drone_steps_per_step": m["active"],
            # the binding roofline of this path is the non-tensor fp64 pipe (arithmetic
            # intensity ~650 flop/B), its peak is measured on the box in this run; the HBM figure the
            # north star also asks for is reported next to it.
            "roofline": {"bound": "fp64 (non-tensor FMA pipe; neither hbm nor tensor binds, SURVEY 8d)",
                         "achieved": fp64_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": fp64_tf / tf.value,
                         "traffic": traffic,
                         "flop_per_drone_step": FLOP_PER_STEP, "kernel": "fly_mlp_f64_kernel", "kernel_ms": kernel_ms,
                         "peak_source": "ld_measure_fp64_peak: DFMA-chain microbenchmark run inside this bench "
                                        "(MEASURED_PEAKS.json has no fp64 figure)",
                         "dfma_per_clk_per_sm": per_clk.value,
                         "executed_fp64_instr_per_drone_step": instr,
                         "pipe_utilisation": (steps_per_launch * instr / (kernel_ms * 1e-3) / (tf.value * 1e12 / 2.0)) if instr else None,
                         "ncu_stamp": stamp_src, "csrc_sha16": csrc_hash(),
                         "note": "achieved = active drone-steps per launch x 850 algorithmic flop / kernel time (CUDA events on the launching "
                                 "stream); pipe_utilisation = executed fp64-pipe thread-instructions per drone-step (ncu source counters, "
                                 "file in ncu_stamp) / measured DFMA issue peak; `traffic` and the instruction count are null when the "
                                 "kernel sources changed after the capture; the co-limiter is dependency latency with one warp per SM "
                                 "sub-partition (brain capacity), DESIGN.md 4"},
            "hbm_roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved_gbs / peaks["hbm_gbs"], "algorithmic_bytes": alg_bytes, "traffic": traffic,
                             "peak_source": peaks_src,
                             "note": "compute-bound by construction: 2,544 algorithmic bytes per 2,000-step episode (SURVEY 8d)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": D * GENOME_BYTES * world,
                    "d2h_bytes_per_step": D * RECORD_BYTES * world,
                    "api": "ld_fly_mlp_f64_host via fly_population_host (one persistent launch; genome rows stream in behind it)",
                    "pipelined": (None if m["pipe_s"] is None else
                                  {"value": m["active"] * args.steps / m["pipe_s"], "unit": UNIT,
                                   "seconds_of_three_runs": m["pipe_runs"],
                                   "api": "ld_fly_mlp_f64_host_submit / _wait via fly_population_host_submit, two submissions in flight: step "
                                          "k+1's host->device copy overlaps step k's kernel; same bytes per step, every step's result read on "
                                          "the host.  For jobs that evaluate independent populations back to back; `value` above is one "
                                          "blocking call at a time"}),
                    "per_rank_seconds": e2e_all,
                    "h2d_gbs_per_rank_concurrent": h2d_all,
                    "limiter": "host->device copy of the genome rows: every rank streams its own pinned buffer through the host at the "
                               "same time; h2d_gbs_per_rank_concurrent is that leg alone (one rank alone reaches ~55 GB/s; on the 8-GPU "
                               "box four GPUs behind one host bridge got 23 GB/s each and the other four 36 GB/s, whatever cores the ranks "
                               "were bound to — profiles/r02_bench_n8*.json); kernel time is `ms_per_step`",
                    "binding": binding},
            "gpu_launches": args.steps * world,
            "clocks": m["clocks"],
        }
        if gens is not None:
            line["generations"] = gens
        if config4 is not None:
            line["config4"], line["config5"] = config4, config5
        if prec != "f64":
            line["note"] = ("float32 FAST MODE: no parity claim (chaotic dynamics, SURVEY 0.4); the roofline objects refer to the "
                            "fp64 pipe and do not apply; the benchmark of record is --precision f64")
        if not args.no_cpu_baseline and world == 1:
            port, _ = port_baseline(args.population, H, PORT_BASELINE_DRONES, D)
            if reference_available() and not args.port_only:
                from oracle.ref_runner import ReferencePool

                pool = ReferencePool()
                pool.fly(population(args.population, D, hi=min(64, 2 * pool.nproc)), H)  # warm the workers
                line["cpu_baseline"], _ = reference_baseline(pool, args.population, H, args.cpu_sample or CPU_BASELINE_DRONES, D)
                pool.close()
                line["cpu_baseline_port"] = port
            else:
                port["sample"] += "; oracle/_ref (the reference's Python files) is not staged on this box"
                line["cpu_baseline"] = port
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
