#!/usr/bin/env python
"""bench.py — drone-steps/s of the population fitness evaluation (fly() path) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--population survivors|random] [--drones-per-gpu D] [--timesteps H]

A "step" is one pass of the hot path over one batch: every drone of the per-GPU population
flies one H-timestep episode (one generation's fitness evaluation).  Workload at N=1 =
BASELINE.json configs[2]: 65,536 synthetic drones, 2,000-step coin episode, the all-survivor
(S) family of SURVEY.md §8d so that active == nominal drone-steps.  N>1 (torchrun, one rank
per GPU): each rank flies its own 65,536-drone shard (weak scaling, drones are independent)
and the fitness records are all-gathered with NCCL inside the timed step.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU algorithm
(the C oracle port — the reference itself is pure Python and does not exist on the GPU box)
on all host cores, on a bounded sample of the same workload.
"""
import argparse
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
FP64_INSTR_PER_STEP = 638.0  # executed fp64-pipe instructions per drone-step (ncu source counters, profiles/r01_h_*)
GENOME_BYTES = 314 * 8
RECORD_BYTES = 32
# dram__bytes_read.sum + dram__bytes_write.sum of fly_mlp_f64_kernel for the default workload
# (65,536 survivors x 2,000 steps), one `ncu --set full` capture: profiles/r01_h_fly_mlp_f64_ncu_full_summary.csv.
# 3.4x the algorithmic bytes ON PURPOSE: the episode is flown in three segments (ld_params.slice_steps, boundaries
# at 500 and 1,500 steps), every drone is parked twice (256 B out + in) and its 2,512-B brain row is read three
# times — 0.4 GB of extra traffic (1.2 % of HBM peak for 7 ms) that removes the 13.5 % wave-quantisation loss of
# 3.46 waves and starts the lanes of a streamed launch evenly.
NCU_TRAFFIC_DEFAULT_WORKLOAD = 521041664 + 39618816


def population(kind, n, seed=2026):
    """The config-3 generators of SURVEY.md §8d (same formulas as oracle/gen_golden.py; the
    golden fixtures are the first rows of exactly these arrays for seed 2026)."""
    rng = np.random.default_rng(seed)
    hover_bias = float(np.log(0.1962 / 0.8038) / 0.4)
    # draw both families in the fixture's order so that prefixes match tests/golden
    g_s = np.zeros((n, 314))
    delta = rng.uniform(-0.010, 0.100, size=n)
    for j in range(4):
        g_s[:, 270 + j * 11 + 10] = hover_bias + delta
    if kind == "survivors":
        return g_s
    cols = [rng.normal(0.0, 1.0 / np.sqrt(r), size=(n, r * c)) for r, c in ((10, 16), (10, 11), (4, 11))]
    return np.concatenate(cols, axis=1)


class ClockSampler:
    """SM clock / throttle-reason sampler running DURING the timed region (NVML, every ~5 ms;
    falls back to `nvidia-smi -lms 20`)."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.thread, self.max_mhz = index, [], False, None, None

    def _loop_nvml(self):
        import pynvml as N

        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM))
        get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.rows.append((time.perf_counter(), float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)),
                                  N.nvmlDeviceGetPowerUsage(h) / 1000.0, int(get_reasons(h))))
            except N.NVMLError:
                pass
            time.sleep(0.005)

    def _loop_smi(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "20"],
                                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        for line in proc.stdout:
            if self.stop_flag:
                break
            f = [x.strip() for x in line.split(",")]
            try:
                self.max_mhz = float(f[1])
                self.rows.append((time.perf_counter(), float(f[0]), float(f[2]), int(f[3], 16)))
            except (ValueError, IndexError):
                pass
        proc.terminate()

    def start(self):
        try:
            import pynvml  # noqa: F401

            target = self._loop_nvml
        except ImportError:
            target = self._loop_smi
        self.thread = threading.Thread(target=target, daemon=True)
        self.thread.start()

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        if not rows:
            return None
        sm = sorted(r[1] for r in rows)
        mask = 0
        for r in rows:
            mask |= r[3]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "power_w_max": max(r[2] for r in rows),
                "samples": len(rows), "reasons": [name for bit, name in self.REASONS if mask & bit]}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline(kind, H, sample, threads=0):
    """The reference's algorithm (C oracle port) on the host cores, bounded sample."""
    from learning_drones_b200._abi import make_params
    from oracle import oracle as O

    g = population(kind, sample)
    p = make_params(H)
    O.fly_mlp(p, g[:64], nthreads=threads)  # warm
    t0 = time.perf_counter()
    rec, _ = O.fly_mlp(p, g, nthreads=threads)
    dt = time.perf_counter() - t0
    steps = int(rec["flight_time"].sum())
    return steps / dt, steps, dt, (threads or O.max_threads())


def workload_config(args, world, active_steps=None):
    """`config` of the JSON line — identical for our arm and the reference arm."""
    D, H = args.drones_per_gpu, args.timesteps
    cfg = {"workload": f"configs[2]: {D}-drone {args.population} population per GPU, {H}-step coin episode, "
                       "batched fitness eval" + (" + NCCL all-gather of fitness records" if world > 1 else ""),
           "population": args.population, "timesteps": H, "drones_per_gpu": D, "global_drones": D * world,
           "nominal_drone_steps_per_step": D * world * H,
           "l2": f"inputs larger than L2 ({D * GENOME_BYTES / 1e6:.0f} MB genome matrix per GPU re-read every step)",
           "parallelism": f"population sharded over {world} GPU(s), no data-path collective; records all-gathered"}
    if active_steps is not None:
        cfg["active_drone_steps_per_step"] = active_steps
    return cfg


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle as O

    cores = O.max_threads()
    sample = args.cpu_sample or 32768
    H = args.timesteps
    for _ in range(args.warmup):
        cpu_baseline(args.population, H, min(sample, 512))
    t0 = time.perf_counter()
    steps = 0
    for _ in range(args.steps):
        v, s, dt, cores = cpu_baseline(args.population, H, sample)
        steps += s
    total = time.perf_counter() - t0
    value = steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} drones x {H} steps per step (first rows of the workload population), "
                                   f"C restatement oracle/ld_oracle.c on {cores} threads; the pure-Python reference "
                                   f"itself runs at ~4.7e3 drone-steps/s/core (BASELINE.md)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--population", default="survivors", choices=["survivors", "random"])
    ap.add_argument("--drones-per-gpu", type=int, default=65536)
    ap.add_argument("--timesteps", type=int, default=2000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="drones in the CPU-baseline sample (default 32768: ~20 CPU-seconds)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="f64 = parity mode (the benchmark); f32 = fast mode, statistical agreement only, never the headline")
    ap.add_argument("--generations", type=int, default=5, help="generations of the full GA loop to time (0 = skip)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

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

    D, H = args.drones_per_gpu, args.timesteps
    params = make_params(H)
    params.refill_batch = int(os.environ.get("LD_REFILL_BATCH", "0"))
    params.slice_steps = int(os.environ.get("LD_SLICE_STEPS", "0"))
    # every rank flies its own shard of the global population [rank*D, (rank+1)*D)
    host_genomes = population(args.population, D * world)[rank * D:(rank + 1) * D]
    pinned = torch.from_numpy(np.ascontiguousarray(host_genomes)).pin_memory()
    genomes = pinned.to(dev)
    records = torch.empty((D, 4), dtype=torch.int64, device=dev)
    counter = torch.empty(2, dtype=torch.int64, device=dev)  # queue head, error flag

    prec = args.precision
    if prec != "f64":
        args.generations = 0

    def step():
        ldpop.fly_population_device(genomes, params, records=records, work_counter=counter, precision=prec)
        if world > 1:
            return allgather_records(records, n_global=D * world)  # fitness all-gather for selection (NCCL/NVLink)
        return records

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- device-resident timing (value): K steps, CUDA events on the launching stream ----------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.05)
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e_all0.record()
    for a, b in ev:
        a.record()
        ldpop.fly_population_device(genomes, params, records=records, work_counter=counter, precision=prec)
        b.record()
        if world > 1:
            allgather_records(records, n_global=D * world)
    e_all1.record()
    barrier()
    t_wall1 = time.perf_counter()
    total_ms = e_all0.elapsed_time(e_all1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    rec_np = ldpop.records_to_numpy(records)
    active_steps_rank = int(rec_np["flight_time"].sum())
    act = torch.tensor([active_steps_rank], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(act)
    active_steps = int(act.item())  # per step, whole job
    value = active_steps * args.steps / (total_ms * 1e-3)

    # ---- end to end through the host C-ABI call (pinned host buffers, copies inside) ----------
    out_pinned = torch.empty((D, 4), dtype=torch.int64).pin_memory()
    out_np = out_pinned.numpy().view(RECORD_DTYPE).reshape(-1)
    host_np = pinned.numpy()
    ldpop.fly_population_host(host_np, params, out=out_np, precision=prec)  # warm (allocates the workspace)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ldpop.fly_population_host(host_np, params, out=out_np, precision=prec)
        _ = int(out_np["reward"].max())  # the step's result is read on the host
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert out_np.tobytes() == rec_np.tobytes(), "host-API result differs from device-API result"
    e2e_value = active_steps * args.steps / e2e_s

    # ---- generations/s: full loop (fly + record all-gather + best tracking + GA step + elitism) ----
    gens = None
    if args.generations > 0:
        from learning_drones_b200.evolve.gpu import GpuEvolution

        ev = GpuEvolution(population(args.population, D * world), seed=2026, timesteps=H)
        ev.step()  # warm (allocations, NCCL)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        hist = ev.run(args.generations)
        g1.record()
        barrier()
        gms = g0.elapsed_time(g1)
        if world > 1:
            tt = torch.tensor([gms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            gms = float(tt.item())
        gens = {"value": args.generations / (gms * 1e-3), "unit": "generations/s", "generations": args.generations,
                "population": D * world, "timesteps": H,
                "active_drone_steps_per_generation": [h["active_steps"] for h in hist],
                "best_score": [h["global_best_score"] for h in hist],
                "note": "one untimed warm-up generation flies the bench population; the timed generations are its GA offspring "
                        "(most crash early), enqueued back to back with one host synchronisation at the end"}
        del ev

    if rank == 0:
        peaks, peaks_src = measured_peaks()
        tf = __import__("ctypes").c_double()
        per_clk = __import__("ctypes").c_double()
        _lib.check(L.ld_measure_fp64_peak(tf, per_clk))
        alg_bytes = D * (GENOME_BYTES + RECORD_BYTES)  # per launch: genome rows in, records out
        achieved_gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
        steps_per_launch = active_steps / world
        fp64_tf = steps_per_launch * FLOP_PER_STEP / (kernel_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": prec, "data": "synthetic",
            "config": workload_config(args, world, active_steps),
            # the binding roofline of this path is the non-tensor fp64 pipe (arithmetic
            # intensity ~650 flop/B), its peak is measured on the box in this run; the HBM figure the
            # north star also asks for is reported next to it.
            "roofline": {"bound": "fp64 (non-tensor FMA pipe; neither hbm nor tensor binds, SURVEY 8d)",
                         "achieved": fp64_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": fp64_tf / tf.value,
                         "traffic": NCU_TRAFFIC_DEFAULT_WORKLOAD if (D, H, args.population) == (65536, 2000, "survivors") else None,
                         "flop_per_drone_step": FLOP_PER_STEP, "kernel": "fly_mlp_f64_kernel", "kernel_ms": kernel_ms,
                         "peak_source": "ld_measure_fp64_peak: DFMA-chain microbenchmark run inside this bench "
                                        "(MEASURED_PEAKS.json has no fp64 figure)",
                         "dfma_per_clk_per_sm": per_clk.value,
                         "executed_fp64_instr_per_drone_step": FP64_INSTR_PER_STEP,
                         "pipe_utilisation": steps_per_launch * FP64_INSTR_PER_STEP / (kernel_ms * 1e-3) / (tf.value * 1e12 / 2.0),
                         "note": "achieved = active drone-steps per launch x 850 algorithmic flop / kernel time; "
                                 "pipe_utilisation = executed fp64-pipe thread-instructions (ncu: 418 DFMA + 116 DMUL + 83 DADD "
                                 "+ 21 DSETP per drone-step, profiles/r01_h_*) / measured DFMA issue peak; the co-limiter is "
                                 "dependency latency with one warp per SM sub-partition (brain capacity), DESIGN.md 4"},
            "hbm_roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved_gbs / peaks["hbm_gbs"], "algorithmic_bytes": alg_bytes,
                             "traffic": NCU_TRAFFIC_DEFAULT_WORKLOAD if (D, H, args.population) == (65536, 2000, "survivors") else None,
                             "peak_source": peaks_src,
                             "note": "compute-bound by construction: 2,544 algorithmic bytes per 2,000-step episode; the measured traffic is "
                                     "3.4x that on purpose (episode flown in three segments: brain row read three times, drone parked "
                                     "twice) — 0.4 GB of idle HBM bandwidth traded for the 13.5 % wave-quantisation loss, DESIGN.md 4"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": D * GENOME_BYTES * world,
                    "d2h_bytes_per_step": D * RECORD_BYTES * world, "api": "ld_fly_mlp_f64_host via fly_population_host (one persistent launch; genome rows stream in behind it)"},
            "gpu_launches": args.steps * world,
            "clocks": clocks,
        }
        if gens is not None:
            line["generations"] = gens
        if prec != "f64":
            line["note"] = ("float32 FAST MODE: no parity claim (chaotic dynamics, SURVEY 0.4); the roofline objects refer to the "
                            "fp64 pipe and do not apply; the benchmark of record is --precision f64")
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample or 32768
            v, s, dt, cores = cpu_baseline(args.population, H, sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} drones x {H} steps ({s} active drone-steps, {dt:.1f} s) of the same "
                                              f"population; C restatement of the reference on {cores} threads"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
