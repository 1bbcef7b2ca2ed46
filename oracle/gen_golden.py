"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

TEST INFRASTRUCTURE.  Runs only in the build container (the reference tree is not
present on the GPU box); the produced fixtures are committed so the parity tests can
run anywhere.  Shims applied: `np.Infinity = np.inf` (removed in
numpy 2.0; reference drone/drone.py:67,178 uses it) and a scratch cwd because every
`Drone()` creates ./evolution/<timestamp>/ (datahandling/data_scraper.py:58-66).

    python oracle/gen_golden.py [--out tests/golden] [--procs 8]
"""
import argparse
import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

REF = os.environ.get("LD_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

LAYER_SHAPES = ((10, 16), (10, 11), (4, 11))
GENOME_LEN = 314
HOVER_BIAS = float(np.log(0.1962 / 0.8038) / 0.4)  # sigmoid(b) = 0.1962 = G/(4 c_f)


def _import_reference():
    np.Infinity = np.inf
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import logging

    logging.disable(logging.CRITICAL)
    with contextlib.redirect_stdout(io.StringIO()):
        from drone.drone import Drone  # noqa
        from flight import fly  # noqa
    return Drone, fly


def split_genome(g):
    out, o = [], 0
    for r, c in LAYER_SHAPES:
        out.append(np.array(g[o : o + r * c], dtype=np.float64).reshape(r, c))
        o += r * c
    return out


def flatten_weights(ws):
    return np.concatenate([np.asarray(w, dtype=np.float64).flatten() for w in ws])


STATE_LEN = 28


def _state_row(d):
    s = d.sensors
    return np.concatenate(
        [
            d.position.ravel(),
            d.velocity.ravel(),
            d.angle.ravel(),
            d.angle_vel.ravel(),
            [x.current_acceleration for x in s],
            [x.last_acceleration for x in s],
            np.asarray(d.thrust, dtype=np.float64).ravel(),
        ]
    ).astype(np.float64)


def ref_fly(args):
    """Fly one genome through the reference. Returns dict of results (+trace if asked)."""
    genome, timesteps, want_trace, kw = args
    Drone, fly = _import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        d = Drone(visualize=False, n_servers=0)
    d.controller.weights = split_genome(genome)
    trace = []
    if want_trace:
        orig_measure, orig_brain = d.measure, d.translate_input_to_thrust

        def measure(wind_speed):
            r = orig_measure(wind_speed=wind_speed)
            trace.append(_state_row(d))
            return r

        def brain(coin_position):
            orig_brain(coin_position)
            trace[-1][24:28] = np.asarray(d.thrust, dtype=np.float64).ravel()

        d.measure, d.translate_input_to_thrust = measure, brain
    err = ""
    try:
        reward, real_time, _ = fly(d, timesteps=timesteps, idx=0, **kw)
    except ValueError as e:  # int(nan)
        reward, real_time, err = d.reward, float("nan"), repr(e)
    return dict(
        reward=int(reward),
        real_time=float(real_time),
        coins=int(d.coins),
        flight_time=int(d.flight_time),
        crashed=bool(d.crashed),
        distance=float(d.distance_to_coin),
        state=_state_row(d),
        trace=np.array(trace) if want_trace else None,
        err=err,
    )


def run_population(genomes, timesteps, pool, want_trace=False, **kw):
    jobs = [(g, timesteps, want_trace, kw) for g in genomes]
    res = pool.map(ref_fly, jobs, chunksize=1) if pool else [ref_fly(j) for j in jobs]
    out = dict(
        genomes=np.asarray(genomes, dtype=np.float64),
        timesteps=np.int64(timesteps),
        reward=np.array([r["reward"] for r in res], dtype=np.int64),
        real_time=np.array([r["real_time"] for r in res]),
        coins=np.array([r["coins"] for r in res], dtype=np.int32),
        flight_time=np.array([r["flight_time"] for r in res], dtype=np.int32),
        crashed=np.array([r["crashed"] for r in res], dtype=np.bool_),
        distance=np.array([r["distance"] for r in res]),
        state=np.array([r["state"] for r in res]),
    )
    if want_trace:
        for i, r in enumerate(res):
            out[f"trace{i}"] = r["trace"]
    return out


# ---- synthetic populations of SURVEY.md §8d, shared with bench.py / tests (kept in sync by
# tests/test_oracle_golden.py which regenerates them from the same formulas) -------------------
def survivors(n, rng):
    """(S) hover family: hidden weights 0, output bias b0 + delta, delta ~ U(-0.010, 0.100) per drone."""
    g = np.zeros((n, GENOME_LEN))
    delta = rng.uniform(-0.010, 0.100, size=n)
    o3 = 160 + 110
    for j in range(4):
        g[:, o3 + j * 11 + 10] = HOVER_BIAS + delta
    return g


def random_brains(n, rng):
    """(R) reference init N(0, 1/sqrt(n_out)) (evolve/brain.py:24)."""
    cols = []
    for r, c in LAYER_SHAPES:
        cols.append(rng.normal(0.0, 1.0 / np.sqrt(r), size=(n, r * c)))
    return np.concatenate(cols, axis=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--procs", type=int, default=os.cpu_count())
    ap.add_argument("--extra", default=None, help="npz with genomes to fly additionally (key 'genomes')")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    out_dir = os.path.abspath(args.out)
    extra = np.load(args.extra)["genomes"] if args.extra else None
    scratch = tempfile.mkdtemp(prefix="ld_golden_")
    os.chdir(scratch)
    Drone, fly = _import_reference()
    import multiprocessing as mp

    pool = mp.get_context("fork").Pool(args.procs) if args.procs > 1 else None
    t0 = time.time()

    # 1. SURVEY §8c: np.random.seed(0); Drone(); fly()  (default 2500 steps), with full trace
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        d = Drone(visualize=False, n_servers=0)
    g0 = flatten_weights(d.controller.weights)[None, :]
    np.savez_compressed(os.path.join(out_dir, "seed0_single.npz"), **run_population(g0, 2500, None, want_trace=True))

    # 2. config 2: np.random.seed(12345); 20x Drone(); one generation
    np.random.seed(12345)
    with contextlib.redirect_stdout(io.StringIO()):
        drones = [Drone(visualize=False, n_servers=0) for _ in range(20)]
    g20 = np.array([flatten_weights(x.controller.weights) for x in drones])
    np.savez_compressed(os.path.join(out_dir, "seed12345_gen0.npz"), **run_population(g20, 2500, pool))

    # 3. hover brain, 2000 steps, trace every step
    gh = np.zeros((1, GENOME_LEN))
    for j in range(4):
        gh[0, 270 + j * 11 + 10] = HOVER_BIAS
    np.savez_compressed(os.path.join(out_dir, "hover_2000.npz"), **run_population(gh, 2000, None, want_trace=True))

    # 4. config 3 subsamples: first rows of the (S) and (R) populations of rng(2026)
    rng = np.random.default_rng(2026)
    gs = survivors(65536, rng)[:96]
    gr = random_brains(65536, rng)[:512]
    np.savez_compressed(os.path.join(out_dir, "survivors_2000.npz"), **run_population(gs, 2000, pool))
    np.savez_compressed(os.path.join(out_dir, "random_2000.npz"), **run_population(gr, 2000, pool))

    # 5. reward flag variants on the seed-12345 brains (punish_time / neither)
    np.savez_compressed(
        os.path.join(out_dir, "seed12345_punish.npz"),
        **run_population(g20, 2500, pool, punish_time=True, reward_time=False),
    )

    # 6. hand-picked interesting brains (coin collectors, long fliers) found with the C oracle
    if extra is not None:
        np.savez_compressed(os.path.join(out_dir, "selected_2500.npz"), **run_population(extra, 2500, pool, want_trace=False))

    print(f"goldens written to {out_dir} in {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
