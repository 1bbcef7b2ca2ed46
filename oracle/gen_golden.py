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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

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


# ---- PID mode (config 1): flight/collect_coins.py ----------------------------------------------
PID_PARAMSETS = ["00", "01", "02", "10", "11", "12", "13", "20", "21", "22"]
PID_COINS = [[0.0, 0.0, 5.0], [25.0, 25.0, 5.0], [-10.0, 35.0, 5.0], [0.0, 0.0, 5.0]]  # collect_coins.py:253-258


def _import_reference_pid():
    _import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        from drone.controller.pid import PID
        from drone.onboard_computer import PIDControlUNnit
        from drone.quadcopter import choose_paramset, parameters
        from drone.quadcopter.physical_model import QuadcopterPhysics
        from drone.sensor.sensor import Sensor, get_positions_and_angles
        from objects.coins import Coin
    # SURVEY §8c second shim: numpy >= 1.24 refuses the ragged array the cascade builds from
    # shape-(1,) setpoints (onboard_computer.py:74-81); numpy 1.17 scalar-converted them.
    if not getattr(PID, "_ld_shim", False):
        orig = PID.set_setpoint

        def set_setpoint(self, point):
            orig(self, float(np.asarray(point).reshape(-1)[0]) if isinstance(point, np.ndarray) else point)

        PID.set_setpoint = set_setpoint
        PID._ld_shim = True
    return PID, PIDControlUNnit, choose_paramset, parameters, QuadcopterPhysics, Sensor, get_positions_and_angles, Coin


def pid_gains(paramset):
    """18 gains in choose_paramset.convert_paramset_2_float order."""
    cp = _import_reference_pid()[2]
    return np.array(cp.convert_paramset_2_float(cp.get_paramset(paramset)), dtype=np.float64)


def ref_fly_pid(args):
    """The main loop of flight/collect_coins.py:125-364 driven through the reference's own PID,
    PIDControlUNnit, QuadcopterPhysics, Sensor and Coin classes (the script itself is a __main__
    block with prints and sleeps; `run_collect_coins_script` runs it verbatim once as a cross-check)."""
    gains, timesteps, want_trace = args
    PID, PIDControlUNnit, cp, P, QuadcopterPhysics, Sensor, get_positions_and_angles, Coin = _import_reference_pid()
    (kp_x, kp_y, kp_z, ki_x, ki_y, ki_z, kd_x, kd_y, kd_z, kp_rol, kp_pit, kp_yaw, ki_rol, ki_pit, ki_yaw, kd_rol, kd_pit, kd_yaw) = [float(v) for v in gains]
    delta_t = 0.01
    lim = [-np.pi / 4, np.pi / 4]
    mk = lambda kp, ki, kd, **kw: PID(kp=kp, ki=ki, kd=kd, timeStep=delta_t, setValue=0, integralRange=2, calculateFlag="signChange", **kw)
    rot_pids = [mk(kp_rol, ki_rol, kd_rol, outputLimitRange=lim), mk(kp_pit, ki_pit, kd_pit, outputLimitRange=lim),
                mk(kp_yaw, ki_yaw, kd_yaw, outputLimitRange=[-np.pi / 6, np.pi / 6])]
    lin_pids = [mk(kp_x, ki_x, kd_x, outputLimitRange=lim), mk(kp_y, ki_y, kd_y, outputLimitRange=lim),
                mk(kp_z, ki_y, kd_z)]  # collect_coins.py:181 uses ki_y for the z PID
    for pid, sv in zip(lin_pids, (10, 10, 10)):
        pid.setValue = sv
    qc = QuadcopterPhysics(mass_center=P.mass_center, mass_motor=P.mass_motor, radius_motor_center=P.radius_motor_center,
                           coef_force=P.coef_force, coef_moment=P.coef_moment, coef_wind=P.coef_wind, gravity=P.gravity,
                           mass_payload=P.mass_payload, x_payload=P.x_payload, y_payload=P.y_payload, I_x=P.I_x, I_y=P.I_y, I_z=P.I_z)
    controller = PIDControlUNnit(np.concatenate([lin_pids, rot_pids]), qc)
    initial_pos = np.array([[0.0, 0.0, 10.0]]).T
    zeros = np.array([[0.0, 0.0, 0.0]]).T
    sensors = [Sensor(delta_t, initial_pos[i, 0], 0.0) for i in range(3)] + [Sensor(delta_t, 0.0, 0.0) for _ in range(3)]
    wind, lab_pos, vel, angle, angle_vel = zeros.copy(), initial_pos, zeros.copy(), zeros.copy(), zeros.copy()
    thrust = np.array([[0, 0, 0, 0]]).T
    coins = [Coin(np.asarray([c]).T, 1, value=1000) for c in PID_COINS]
    target_index = new_target_index = 0
    score = timesteps
    steps, pick_steps, trace, finished = 0, [], [], False
    for time_i in range(timesteps):
        score -= 1
        steps += 1
        if coins[target_index].is_in_vicinity(lab_pos):
            score += coins[target_index].value
            new_target_index = target_index + 1
            pick_steps.append(time_i)
        if new_target_index >= len(coins):
            finished = True
            break
        rot_targets = np.zeros((3, 1))
        lin_targets = coins[new_target_index].position
        [rot_pids[i].set_setpoint(rot_targets[i, 0]) for i in range(3)]
        [lin_pids[i].set_setpoint(lin_targets[i, 0]) for i in range(3)]
        target_index = new_target_index
        lin_acc, rot_acc = qc.calculate_accelerations(angle, wind, thrust, lin_acc_drone_2_lab=True)
        [sensors[i].measure_acceleration(lin_acc[i, 0]) for i in range(3)]
        [sensors[i].measure_acceleration(rot_acc[i - 3, 0]) for i in range(3, 6)]
        lab_pos, vel, angle, angle_vel = get_positions_and_angles(sensors)
        thrust = controller.translate_input_to_thrust(lab_pos, vel, angle, angle_vel, angle, lin_pids, rot_pids)
        if want_trace:
            trace.append(np.concatenate([lab_pos.ravel(), vel.ravel(), angle.ravel(), angle_vel.ravel(),
                                         [s.current_acceleration for s in sensors], [s.last_acceleration for s in sensors],
                                         np.asarray(thrust, dtype=np.float64).ravel()]))
    state = np.concatenate([lab_pos.ravel(), vel.ravel(), angle.ravel(), angle_vel.ravel(),
                            [s.current_acceleration for s in sensors], [s.last_acceleration for s in sensors],
                            np.asarray(thrust, dtype=np.float64).ravel()]).astype(np.float64)
    return dict(score=int(score), steps=steps, coins=len(pick_steps), finished=finished, pick_steps=pick_steps, state=state,
                trace=np.array(trace) if want_trace else None)


def run_collect_coins_script(paramset="11"):
    """Run the reference's flight/collect_coins.py verbatim (headless) and return its printed Final Score."""
    import re
    import runpy
    import time as _time

    _import_reference_pid()
    out = io.StringIO()
    argv, sleep = sys.argv, _time.sleep
    sys.argv = ["collect_coins.py", "65432", "0", "0", paramset]
    _time.sleep = lambda s: None
    try:
        with contextlib.redirect_stdout(out):
            runpy.run_path(os.path.join(REF, "flight", "collect_coins.py"), run_name="__main__")
    except (SystemExit, AttributeError):  # dh.finish() touches self.conn, undefined when not visualising
        pass
    finally:
        sys.argv, _time.sleep = argv, sleep
    m = re.search(r"Final Score: (-?\d+)", out.getvalue())
    return int(m.group(1)) if m else None


def main_pid(out_dir, pool):
    rng = np.random.default_rng(11)
    gains = [pid_gains(ps) for ps in PID_PARAMSETS]
    # plus perturbed copies of paramset 11 (a "batch of gain sets")
    base = pid_gains("11")
    gains += [base * rng.uniform(0.7, 1.3, size=18) for _ in range(54)]
    gains = np.array(gains)
    jobs = [(g, 3500, i == 4) for i, g in enumerate(gains)]
    res = pool.map(ref_fly_pid, jobs, chunksize=1) if pool else [ref_fly_pid(j) for j in jobs]
    script_score = run_collect_coins_script("11")
    pick = np.full((len(res), 4), -1, dtype=np.int32)
    for i, r in enumerate(res):
        pick[i, : len(r["pick_steps"])] = r["pick_steps"]
    np.savez_compressed(
        os.path.join(out_dir, "pid_collect_coins.npz"), gains=gains, paramsets=np.array(PID_PARAMSETS), timesteps=np.int64(3500),
        score=np.array([r["score"] for r in res], dtype=np.int64), steps=np.array([r["steps"] for r in res], dtype=np.int32),
        coins=np.array([r["coins"] for r in res], dtype=np.int32), finished=np.array([r["finished"] for r in res]),
        pick_steps=pick, state=np.array([r["state"] for r in res]), trace11=res[4]["trace"],
        script_final_score_11=np.int64(-1 if script_score is None else script_score))
    print("pid goldens:", [(ps, r["score"], r["steps"]) for ps, r in zip(PID_PARAMSETS, res)], "script(11) =", script_score)


def evolution_loop(Drone, fly_generation, util, seed, n_drones=20, n_generations=6, mutation_rate=0.05, p_tournament=0.6,
                   p_crossover=0.05):
    """The generation loop of evolve/evolution.py:40-107 (the script itself cannot be imported: it needs
    jsonpickle and runs under __main__).  `fly_generation(drones)` -> [(score, time, idx), ...]."""
    from copy import deepcopy

    np.random.seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        drones = [Drone(visualize=False, n_servers=0) for _ in range(n_drones)]
    global_best_score, global_best_drone, history = -np.inf, None, []
    for gen in range(n_generations):
        scores, times = [0] * n_drones, [0] * n_drones
        for score, flight_time, idx in fly_generation(drones):      # :64-73
            scores[idx], times[idx] = score, flight_time
        best_index = int(np.argmax(scores))                          # :77-78
        if scores[best_index] > global_best_score:                   # :80-83
            global_best_score = scores[best_index]
            global_best_drone = deepcopy(drones[best_index])
        new_population = util.selection_and_crossover(drones, p_tournament, p_crossover)  # :89
        for d in new_population:                                     # :91-93
            d.reset()
            d.mutate(mutation_rate)
        drones = new_population                                      # :105-107
        drones[0] = global_best_drone
        drones[0].reset()
        history.append(dict(scores=list(scores), times=list(times), best_index=best_index, global_best=global_best_score))
    return history


def main_evolution(out_dir):
    Drone, fly = _import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        import evolve.util as rutil
    hist = evolution_loop(Drone, lambda drones: [fly(d, idx=i) for i, d in enumerate(drones)], rutil, seed=12345)
    np.savez_compressed(os.path.join(out_dir, "evolution_seed12345.npz"),
                        scores=np.array([h["scores"] for h in hist], dtype=np.int64), times=np.array([h["times"] for h in hist]),
                        best_index=np.array([h["best_index"] for h in hist]), global_best=np.array([h["global_best"] for h in hist]))
    print("evolution fixture:", [h["global_best"] for h in hist])


def main_large(out_dir, pool, n=8192):
    """Reference outputs for rows [512, 512+n) of the config-3 random population (seed 2026).  Only the
    results are stored — the genomes are regenerated from the seed by `random_brains` (guarded by
    tests/test_oracle_golden.py::test_population_generators_match_fixture)."""
    rng = np.random.default_rng(2026)
    survivors(65536, rng)
    g = random_brains(65536, rng)[512:512 + n]
    r = run_population(g, 2000, pool)
    np.savez_compressed(os.path.join(out_dir, "random_2000_large.npz"), first_row=np.int64(512), n=np.int64(n), timesteps=r["timesteps"],
                        reward=r["reward"], coins=r["coins"], flight_time=r["flight_time"], crashed=r["crashed"],
                        distance=r["distance"], position=r["state"][:, :3], checksum=np.float64(g.sum()))
    print("large fixture:", n, "drones; crashed", int(r["crashed"].sum()), "coins", int(r["coins"].sum()))


def main_brain(out_dir):
    """Reference NeuralNetwork (evolve/brain.py:10-76) for several layer_specs incl. its own default {1: 8, 2: 10, 3: 4}:
    weights from np.random.seed, feed_forward / relu / sigmoid outputs."""
    _import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        from evolve.brain import NeuralNetwork
    out = {}
    specs = {"default": (None, 15), "drone": ({1: 10, 2: 10, 3: 4}, 15), "small": ({1: 5, 2: 3}, 6), "deep": ({1: 12, 2: 7, 3: 9, 4: 2}, 20)}
    for name, (spec, n_in) in specs.items():
        np.random.seed(len(name))
        nn = NeuralNetwork(n_inputs=n_in) if spec is None else NeuralNetwork(layer_spec=dict(spec), n_inputs=n_in)
        xs = np.random.standard_normal((16, n_in, 1)) * np.random.uniform(0.1, 20.0, size=(16, 1, 1))
        ys = np.array([nn.feed_forward(x) for x in xs])
        out[f"{name}_sizes"] = np.array([n_in] + [w.shape[0] for w in nn.weights], dtype=np.int32)
        out[f"{name}_weights"] = np.concatenate([w.flatten() for w in nn.weights])
        out[f"{name}_inputs"], out[f"{name}_outputs"] = xs, ys
    nn = NeuralNetwork()
    v = np.linspace(-30.0, 30.0, 41).reshape(-1, 1)
    out["act_x"], out["act_relu"], out["act_relu_t"], out["act_sigmoid"] = v, nn.relu(v), nn.relu(v, thresh=2.5, alpha=0.5), nn.sigmoid(v)
    # thrust for a full state (brain.py:56-76)
    np.random.seed(3)
    nn = NeuralNetwork(layer_spec={1: 10, 2: 10, 3: 4})
    args = [np.random.standard_normal((3, 1)) * s for s in (10, 3, 1, 5, 10)]
    out["thrust_weights"] = np.concatenate([w.flatten() for w in nn.weights])
    out["thrust_args"], out["thrust_out"] = np.array(args), nn.translate_input_to_thrust(*args)
    np.savez_compressed(os.path.join(out_dir, "brain_forward.npz"), **out)
    print("brain fixture:", {k: v.shape for k, v in out.items() if k.endswith("_outputs")})


def main_starts(out_dir):
    """fly() on Drones that are NOT in the reset state (flight/__init__.py:35-45 continues from whatever the Drone
    holds): constructor arguments (drone/drone.py:20-76), and a second flight without reset()."""
    Drone, fly = _import_reference()
    rng = np.random.default_rng(99)
    genomes = np.concatenate([survivors(6, rng), random_brains(10, rng)])
    n = len(genomes)
    init = dict(pos=rng.uniform(-3, 3, size=(n, 3)) + [0, 0, 20], vel=rng.uniform(-1, 1, size=(n, 3)),
                angle=rng.uniform(-0.2, 0.2, size=(n, 3)), angle_vel=rng.uniform(-0.5, 0.5, size=(n, 3)))
    first, second = [], []
    for i, g in enumerate(genomes):
        with contextlib.redirect_stdout(io.StringIO()):
            d = Drone(visualize=False, n_servers=0, initial_pos=init["pos"][i].reshape(3, 1).copy(), initial_vel=init["vel"][i].reshape(3, 1).copy(),
                      initial_angle=init["angle"][i].reshape(3, 1).copy(), initial_angle_vel=init["angle_vel"][i].reshape(3, 1).copy())
        d.controller.weights = split_genome(g)
        r1 = fly(d, timesteps=300, idx=i)
        first.append((r1[0], r1[1], d.coins, d.flight_time, d.crashed, d.distance_to_coin, _state_row(d)))
        r2 = fly(d, timesteps=200, idx=i)  # no reset: continues, counters accumulate, the coin schedule restarts
        second.append((r2[0], r2[1], d.coins, d.flight_time, d.crashed, d.distance_to_coin, _state_row(d)))
    pack = lambda rows, k: np.array([r[k] for r in rows])
    out = dict(genomes=genomes, **{f"init_{k}": v for k, v in init.items()})
    for name, rows in (("first", first), ("second", second)):
        out.update({f"{name}_reward": pack(rows, 0).astype(np.int64), f"{name}_real_time": pack(rows, 1), f"{name}_coins": pack(rows, 2),
                    f"{name}_flight_time": pack(rows, 3), f"{name}_crashed": pack(rows, 4), f"{name}_distance": pack(rows, 5),
                    f"{name}_state": pack(rows, 6)})
    np.savez_compressed(os.path.join(out_dir, "start_states.npz"), **out)
    print("start-state fixture: first", out["first_reward"].tolist(), "second", out["second_reward"].tolist())


if __name__ == "__main__" and "--starts-only" in sys.argv:
    os.chdir(tempfile.mkdtemp(prefix="ld_golden_"))
    main_starts(os.path.join(ROOT, "tests", "golden"))
    sys.exit(0)


def main_checkpoint(out_dir):
    """evolution/best_drone_<t>.json as the reference writes it (evolve/evolution.py:80-87): the seed-0 Drone after its
    flight, deep-copied, jsonpickle-encoded (restated pickler, oracle/mini_jsonpickle.py — jsonpickle itself is not
    installed here) and json.dump-ed a second time."""
    import json
    from copy import deepcopy

    from oracle import mini_jsonpickle as mj

    Drone, fly = _import_reference()
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        d = Drone(visualize=False, n_servers=0)
    fly(d, idx=0)
    best = deepcopy(d)
    with open(os.path.join(out_dir, "reference_best_drone.json"), "w") as f:
        json.dump(mj.encode(best), f)
    print("checkpoint fixture: reward", best.reward, "flight_time", best.flight_time)


if __name__ == "__main__" and "--checkpoint-only" in sys.argv:
    os.chdir(tempfile.mkdtemp(prefix="ld_golden_"))
    main_checkpoint(os.path.join(ROOT, "tests", "golden"))
    sys.exit(0)


if __name__ == "__main__" and "--brain-only" in sys.argv:
    os.chdir(tempfile.mkdtemp(prefix="ld_golden_"))
    main_brain(os.path.join(ROOT, "tests", "golden"))
    sys.exit(0)


if __name__ == "__main__" and "--evolution-only" in sys.argv:
    sys.argv.remove("--evolution-only")
    os.chdir(tempfile.mkdtemp(prefix="ld_golden_"))
    main_evolution(os.path.join(ROOT, "tests", "golden"))
    sys.exit(0)


if __name__ == "__main__" and "--large-only" in sys.argv:
    sys.argv.remove("--large-only")
    _out = os.path.join(ROOT, "tests", "golden")
    os.chdir(tempfile.mkdtemp(prefix="ld_golden_"))
    import multiprocessing as _mp

    _import_reference()
    main_large(_out, _mp.get_context("fork").Pool(os.cpu_count()))
    sys.exit(0)


if __name__ == "__main__" and "--pid-only" in sys.argv:
    sys.argv.remove("--pid-only")
    _out = os.path.join(ROOT, "tests", "golden")
    _scratch = tempfile.mkdtemp(prefix="ld_golden_")
    os.chdir(_scratch)
    os.makedirs("translation/triple/11", exist_ok=True)
    import multiprocessing as _mp

    main_pid(_out, _mp.get_context("fork").Pool(os.cpu_count()))
    sys.exit(0)


if __name__ == "__main__":
    main()
