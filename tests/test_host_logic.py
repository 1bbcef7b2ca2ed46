"""CPU tests of the host side: C-ABI surface, ld_params derivation, reference-API mirror.

Where the reference tree is available (build container) the mirror classes are compared with
the reference's own classes under identical seeds; on the GPU box those tests skip."""
import contextlib
import ctypes as C
import io
import math
import os
import re
import sys

import numpy as np
import pytest

from learning_drones_b200 import _abi, _lib
from learning_drones_b200._abi import GENOME_LEN, LAYER_SHAPES, LdParams, LdRecord, make_params, paramset_to_pid_gains, sqrt_threshold

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "evolve")), reason="reference tree not present")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ldrones.h")).read()
    declared = set(re.findall(r"^(?:int|int64_t|void|const char\*)\s+(ld_[a-z0-9_]+)\s*\(", hdr, re.M))
    L = _lib.load()  # loads without a GPU; raises if the .so is missing
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.ld_abi_version() == _abi.LD_ABI_VERSION


def test_header_constants_match_the_python_binding():
    hdr = open(os.path.join(ROOT, "include", "ldrones.h")).read()
    macro = lambda name: int(re.search(rf"#define\s+{name}\s+\(?(-?\d+)\)?", hdr).group(1))
    assert macro("LD_ABI_VERSION") == _abi.LD_ABI_VERSION
    assert macro("LD_SLICE_STAGGER") == _abi.SLICE_STAGGER
    assert macro("LD_STATE_LEN") == _abi.STATE_LEN and macro("LD_PID_GAINS") == _abi.PID_GAINS
    assert macro("LD_MAX_PEERS") == _abi.MAX_PEERS and macro("LD_PID_MAX_COINS") == _abi.PID_MAX_COINS
    assert macro("LD_HOST_PIPELINE_DEPTH") == 2 and macro("LD_PROFILE_SLOTS") >= 16
    from learning_drones_b200.population import PROFILE_SLOTS

    assert len(PROFILE_SLOTS) <= macro("LD_PROFILE_SLOTS")


def test_struct_layout_matches_header():
    # sizes implied by include/ldrones.h (doubles 8-aligned, explicit pads)
    assert C.sizeof(LdRecord) == 32
    assert C.sizeof(LdParams) % 8 == 0
    p = make_params()
    assert LdParams.dt.offset == 8 * 10 + 4 * 8  # 10 doubles, then 2+3+3 floats
    assert p.timesteps == 2500 and p.reward_time == 1 and p.punish_time == 0


def test_params_follow_reference_constants():
    p = make_params(2000)
    G = 9.81 * (4 * 0.050 + 0.200)
    assert p.neg_G32 == np.float32(-G) and p.m_total32 == np.float32(0.2 + 4 * 0.05 + 0)
    assert p.sqrt2_r_m == float(np.sqrt(2) * 0.1)
    assert list(p.inv_I) == [1 / 2160e-6, 1 / 2160e-6, 1 / 4210e-6]
    assert list(p.wind_cw32) == [0.0, 0.0, 0.0] and list(p.moments_payload) == [0.0, 0.0, 0.0]
    assert list(p.init_pos) == [0.0, 0.0, 10.0] and list(p.init_thrust) == [0.0] * 4
    assert (p.reward_time_passed, p.reward_crashed, p.reward_unstable, p.reward_coin_distance) == (2, -3000, -1, -100)
    assert p.angle_limit == np.pi / 4 and p.angle_vel_limit == 20 and (p.z_min, p.z_max) == (0.0, 80.0)
    assert list(p.coin0_pos) == [5.0, 5.0, 10.0] and list(p.coin_step) == [5.0, 2.5, 0.0]
    assert (p.coin0_value, p.coinN_value, p.coin0_radius, p.coinN_radius) == (2000, 1000, 2.0, 1.0)
    w = make_params(10, wind_speed=[0.5, -2.0, 0.25], mass_payload=0.05, x_payload=0.01, y_payload=-0.02)
    assert list(w.wind_cw32) == [float(np.float32(v) * np.float32(0.1)) for v in (0.5, -2.0, 0.25)]
    assert w.neg_G32 == np.float32(-(G + 9.81 * 0.05))
    with pytest.raises(ValueError):
        make_params(-1)


@pytest.mark.parametrize("r", [1.0, 2.0, 1.25, 1.5, 0.1, 3.3, 1e-3, 7.0])
def test_sqrt_threshold_is_exact(r):
    thr = sqrt_threshold(r)
    below = math.nextafter(thr, 0.0)
    assert math.sqrt(thr) >= r and math.sqrt(below) < r
    # the reference's test `norm < radius` and the kernel's `norm^2 < thr` agree around the boundary
    for s in (below, thr, math.nextafter(thr, math.inf), r * r):
        assert (math.sqrt(s) < r) == (s < thr)
    assert sqrt_threshold(1.0) == 1.0 and sqrt_threshold(2.0) == 4.0


def test_genome_roundtrip_and_flatten_order():
    from learning_drones_b200.evolve.brain import NeuralNetwork

    np.random.seed(3)
    nn = NeuralNetwork(layer_spec={1: 10, 2: 10, 3: 4})
    assert tuple(w.shape for w in nn.weights) == LAYER_SHAPES
    g = nn.genome()
    assert g.shape == (GENOME_LEN,)
    # layer.flatten() order of evolve/util.py:58-59, bias last column
    assert g[15] == nn.weights[0][0, 15] and g[16] == nn.weights[0][1, 0] and g[160 + 10] == nn.weights[1][0, 10]
    assert g[270 + 11 * 3 + 10] == nn.weights[2][3, 10]
    nn2 = NeuralNetwork(layer_spec={1: 10, 2: 10, 3: 4})
    nn2.set_genome(g)
    assert all(np.array_equal(a, b) for a, b in zip(nn.weights, nn2.weights))
    with pytest.raises(ValueError):
        NeuralNetwork().genome()  # default 15-8-10-4 shape is not the Drone's


def test_drone_reset_and_guards():
    from learning_drones_b200.drone.drone import Drone

    d = Drone(visualize=False, n_servers=0)
    assert d.position.shape == (3, 1) and d.thrust.shape == (4, 1) and d.distance_to_coin == np.inf
    assert len(d.sensors) == 6 and d.sensors[2].previous_pos == 10.0
    # constructor arguments set the state the first flight starts from (reference drone/drone.py:55-76) ...
    e = Drone(initial_pos=np.array([[1.0], [2.0], [30.0]]), initial_angle_vel=np.array([[0.1], [0.2], [0.3]]))
    assert e.position.ravel().tolist() == [1.0, 2.0, 30.0] and e.sensors[5].previous_vel == 0.3 and e.sensors[1].previous_pos == 2.0
    row, reward, coins, flight_time, crashed = e.start_state()
    assert row[:3].tolist() == [1.0, 2.0, 30.0] and row[9:12].tolist() == [0.1, 0.2, 0.3] and (reward, coins, flight_time, crashed) == (0, 0, 0, False)
    # ... a Drone that already flew continues from what it holds (flight/__init__.py:35-45) ...
    d.reward, d.flight_time = 5, 3
    assert d.start_state()[1:] == (5, 0, 3, False)
    d.position = d.position + 1.0  # attributes edited behind the sensors' back: refused, not silently flown
    with pytest.raises(ValueError):
        d.start_state()
    # ... and reset() always goes back to the module-level initial state (:177-218)
    d.reset()
    e.reset()
    assert d.reward == 0 and d.flight_time == 0 and not d.crashed
    assert e.position.ravel().tolist() == [0.0, 0.0, 10.0] and e.sensors[5].previous_vel == 0.0
    with pytest.raises(NotImplementedError):
        Drone(visualize=True, n_servers=1)


def _ref_modules():
    np.Infinity = np.inf
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import logging

    logging.disable(logging.CRITICAL)
    with contextlib.redirect_stdout(io.StringIO()):
        import drone.drone as rdrone
        import evolve.brain as rbrain
        import evolve.util as rutil
    return rdrone, rbrain, rutil


@needs_ref
def test_mirror_classes_match_reference_under_same_seed(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)  # reference Drone() creates ./evolution/<timestamp>/
    rdrone, rbrain, rutil = _ref_modules()
    from learning_drones_b200.drone.drone import Drone
    from learning_drones_b200.evolve import util
    from learning_drones_b200.evolve.brain import NeuralNetwork

    # brain init: same draws from the global numpy stream
    np.random.seed(11)
    a = rbrain.NeuralNetwork(layer_spec={1: 10, 2: 10, 3: 4})
    np.random.seed(11)
    b = NeuralNetwork(layer_spec={1: 10, 2: 10, 3: 4})
    assert all(np.array_equal(x, y) for x, y in zip(a.weights, b.weights))
    # the reference's own default layer_spec {1: 8, 2: 10, 3: 4} (evolve/brain.py:10)
    np.random.seed(12)
    a = rbrain.NeuralNetwork()
    np.random.seed(12)
    b = NeuralNetwork()
    assert [w.shape for w in b.weights] == [(8, 16), (10, 9), (4, 11)] and b.n_layers == a.n_layers == 3
    assert all(np.array_equal(x, y) for x, y in zip(a.weights, b.weights))
    # (feed_forward itself runs on the GPU: tests/test_gpu_parity.py::test_brain_forward_matches_reference_fixture)

    def population(cls, seed, n=8):
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            ds = [cls(visualize=False, n_servers=0) for _ in range(n)]
        for i, d in enumerate(ds):
            d.reward = int((i * 7919) % 13) - 5
        return ds

    ref_pop, our_pop = population(rdrone.Drone, 5), population(Drone, 5)
    assert all(np.array_equal(x, y) for r, o in zip(ref_pop, our_pop) for x, y in zip(r.controller.weights, o.controller.weights))
    # selection + crossover + mutate: identical RNG consumption and identical children
    np.random.seed(77)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_new = rutil.selection_and_crossover(ref_pop, 0.6, 0.5)
    for d in ref_new:
        d.mutate(0.05)
    ref_state = np.random.random_sample()
    np.random.seed(77)
    our_new = util.selection_and_crossover(our_pop, 0.6, 0.5)
    for d in our_new:
        d.mutate(0.05)
    assert np.random.random_sample() == ref_state
    for r, o in zip(ref_new, our_new):
        assert all(np.array_equal(x, y) for x, y in zip(r.controller.weights, o.controller.weights))
    assert util.get_best_of_generation([3, 9, 9, -1]) == rutil.get_best_of_generation([3, 9, 9, -1]) == (1, 9)
    util.compare_weights(our_new[0], our_new[0])
    with pytest.raises(AssertionError):
        util.compare_weights(our_new[0], our_new[1])


def test_bench_population_matches_fixture_generators(golden):
    sys.path.insert(0, ROOT)
    import bench

    assert np.array_equal(bench.population("survivors", 65536)[:96], golden("survivors_2000")["genomes"])
    assert np.array_equal(bench.population("random", 65536)[:512], golden("random_2000")["genomes"])


def test_trace_dump_in_datahandler_layout(tmp_path, golden):
    """a fly() trace written in the layout of the reference's DataHandler."""
    from learning_drones_b200.datahandling import CSV_HEADER, coin_schedule, save_trace

    g = golden("hover_2000")
    trace = g["trace0"][:300]
    res = save_trace(str(tmp_path / "run"), trace)
    assert res.shape == (301, 25) and len(CSV_HEADER.split(",")) == 25
    for name, cols in (("time", 1), ("thrust_values", 4), ("rotation", 3), ("translation", 3), ("wind", 2), ("pid", 6), ("setpoints", 6)):
        a = np.load(tmp_path / "run" / f"{name}.npy")
        assert a.dtype == np.float32 and a.shape == (301, cols) and not a[0].any()
    assert np.allclose(np.load(tmp_path / "run" / "translation.npy")[1:], trace[:, :3].astype(np.float32))
    assert open(tmp_path / "run" / "Results.csv").readline().startswith("# Time, Roll, Pitch")
    sched = coin_schedule(trace)
    assert (sched == [5.0, 5.0, 10.0]).all()  # the hover drone never reaches the first coin
    # a path through the first coin switches the set-point to the next one (radius 2 then 1)
    path = np.zeros((4, 28))
    path[:, :3] = [[4.9, 4.9, 10.0], [4.9, 4.9, 10.0], [10.0, 7.5, 10.0], [0, 0, 0]]
    assert coin_schedule(path).tolist() == [[5, 5, 10], [10, 7.5, 10], [10, 7.5, 10], [15, 10, 10]]


def test_abi_argument_validation_without_gpu():
    """Entry points reject bad arguments with LD_ERR_INVALID_ARG (-1) and a message before touching CUDA."""
    from learning_drones_b200._abi import LdGaParams, make_ga_params, make_peer_table, make_pid_params

    L = _lib.load()
    p = make_params(10)
    assert L.ld_fly_mlp_f64_device(None, None, 0, None, None, None, None) == -1
    assert b"params" in L.ld_last_error()
    assert L.ld_fly_mlp_f64_device(C.byref(p), None, -1, None, None, None, None) == -1
    assert L.ld_fly_mlp_f64_device(C.byref(p), None, 5, None, None, None, None) == -1  # NULL buffers
    assert L.ld_fly_mlp_f64_device(C.byref(p), 8, 5, 16, None, 32, None) == -1        # misaligned genomes
    assert b"16-byte" in L.ld_last_error()
    bad = make_params(10)
    bad.dt = 0.0
    assert L.ld_fly_mlp_f64_host(C.byref(bad), None, 1, None, None) == -1
    long = make_params((1 << 27) + 1)
    assert L.ld_fly_mlp_f64_host(C.byref(long), None, 1, None, None) == -1 and b"2^27" in L.ld_last_error()
    assert L.ld_fly_mlp_f64_host(C.byref(p), None, 0, None, None) == 0  # empty population is a no-op
    assert L.ld_track_best_device(None, None, None, 0, None, None, None, None) == -1
    q = make_pid_params()
    q.n_coins = 99
    assert L.ld_fly_pid_f64_device(C.byref(p), C.byref(q), 16, 4, 16, None, None) == -1
    ga = make_ga_params(1, 0)
    assert L.ld_ga_step_device(C.byref(ga), 16, 16, 7, 16, 0, 6, None) == -1   # odd population
    assert L.ld_ga_step_device(C.byref(ga), 16, 16, 8, 16, 1, 7, None) == -1   # odd shard bounds
    assert b"even" in L.ld_last_error()
    t = make_peer_table([4096, 8192], 4)
    assert L.ld_ga_step_p2p_device(C.byref(ga), C.byref(t), 16, 10, 16, 0, 4, None) == -1  # n != world * shard_rows
    with pytest.raises(ValueError):
        make_peer_table([], 4)
    with pytest.raises(_lib.LdError):
        _lib.check(-1)


def test_properties_with_hypothesis():
    """Property tests of the host-side arithmetic helpers."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from learning_drones_b200.distributed import shard_bounds

    @settings(max_examples=300, deadline=None)
    @given(st.floats(min_value=1e-6, max_value=1e6, allow_nan=False))
    def threshold(r):
        thr = sqrt_threshold(r)
        assert math.sqrt(thr) >= r > math.sqrt(math.nextafter(thr, 0.0))

    @settings(max_examples=300, deadline=None)
    @given(st.integers(min_value=0, max_value=10**7), st.integers(min_value=1, max_value=64))
    def shards(n, world):
        spans = [shard_bounds(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] and 0 <= a[1] - a[0] - (b[1] - b[0]) <= 1 for a, b in zip(spans, spans[1:]))

    @settings(max_examples=100, deadline=None)
    @given(st.lists(st.floats(min_value=-10, max_value=10, allow_nan=False), min_size=18, max_size=18))
    def wiring(g18):
        w = paramset_to_pid_gains(np.array(g18))
        assert w.shape == (6, 3) and w[2, 1] == g18[4] and w[0, 0] == g18[0] and w[5, 2] == g18[17]

    threshold()
    shards()
    wiring()


def test_best_drone_checkpoint_round_trip(tmp_path):
    """persist the best drone (evolution.py:85-87) and load it back (fly_best_drone.py:11-14), in the
    reference's double-encoded jsonpickle format (default) and in the plain one; no RNG draws while loading."""
    import json

    from learning_drones_b200 import checkpoint
    from learning_drones_b200.drone.drone import Drone

    np.random.seed(5)
    d = Drone()
    d.reward, d.coins, d.flight_time, d.crashed, d.distance_to_coin = -4077, 3, 170, True, 1.0 / 3.0
    for fmt in ("jsonpickle", "plain"):
        path = checkpoint.save_drone(d, tmp_path / f"best_drone_{fmt}.json", fmt=fmt)
        rng_before = np.random.get_state()[1].copy()
        e = checkpoint.load_drone(path)
        assert np.array_equal(np.random.get_state()[1], rng_before)  # loading does not consume numpy's global stream
        assert all(np.array_equal(a, b) for a, b in zip(d.controller.weights, e.controller.weights))  # bit for bit
        assert np.array_equal(d.controller.genome(), e.controller.genome())
        assert (e.reward, e.coins, e.flight_time, e.crashed, e.distance_to_coin) == (-4077, 3, 170, True, 1.0 / 3.0)
        assert e.controller.layer_spec == d.controller.layer_spec and e.controller.n_layers == 3
    with open(tmp_path / "best_drone_jsonpickle.json") as f:
        text = json.load(f)
    assert isinstance(text, str)  # evolution.py:85-87 dumps the jsonpickle TEXT: a JSON string, not an object
    doc = json.loads(text)
    assert doc["py/object"] == "drone.drone.Drone" and doc["controller"]["py/object"] == "evolve.brain.NeuralNetwork"
    assert list(doc)[1:] == ["quadcopter", "dh", "sensors", "controller", "distance_to_coin", "reward", "coins", "flight_time",
                             "crashed", "position", "velocity", "angle", "angle_vel", "thrust"]  # Drone.__init__ order
    assert doc["controller"]["layer_spec"] == {"1": 10, "2": 10, "3": 4, "0": 15}
    d.distance_to_coin = np.inf  # a Drone that never flew (drone.py:67): json.dump writes Infinity like the reference's file
    assert checkpoint.load_drone(checkpoint.save_drone(d, tmp_path / "inf.json")).distance_to_coin == np.inf
    with pytest.raises(ValueError):
        checkpoint.save_drone(d, tmp_path / "x.json", fmt="pickle")


def test_reads_reference_checkpoint_fixture(golden):
    """tests/golden/reference_best_drone.json = the reference's seed-0 Drone after its flight, encoded the way
    jsonpickle encodes it (py/object tags, ndarray / numpy-scalar __reduce__ forms, py/id back-references to the shared
    dtype; oracle/gen_golden.py --checkpoint-only)."""
    from learning_drones_b200 import checkpoint

    d = checkpoint.load_drone(os.path.join(ROOT, "tests", "golden", "reference_best_drone.json"))
    g = golden("seed0_single")
    assert np.array_equal(d.controller.genome(), g["genomes"][0])  # every weight bit for bit
    assert (d.reward, d.flight_time, d.coins, d.crashed) == (-4077, 170, 0, True) and d.distance_to_coin == 11.208182019123052
    assert d.controller.layer_spec == {1: 10, 2: 10, 3: 4, 0: 15}
    assert d.position.ravel().tolist() == [0.0, 0.0, 10.0]  # in the reset state, like fly_best_drone.py:14 leaves it


@needs_ref
def test_our_checkpoint_replays_in_the_reference(tmp_path, monkeypatch):
    """A file written by save_drone decodes into the REFERENCE's own Drone class (through the restated jsonpickle
    unpickler, oracle/mini_jsonpickle.py) and re-flies there with the known result — fly_best_drone.py:11-20."""
    import json

    monkeypatch.chdir(tmp_path)
    rdrone, rbrain, rutil = _ref_modules()
    with contextlib.redirect_stdout(io.StringIO()):
        from flight import fly as ref_fly
        import drone.sensor.sensor as rsensor
    from learning_drones_b200 import checkpoint
    from learning_drones_b200.drone.drone import Drone
    from oracle import mini_jsonpickle as mj

    np.random.seed(0)
    ours = Drone()
    path = checkpoint.save_drone(ours, tmp_path / "best_drone_1.json")
    with open(path) as f:
        jsonstring = json.load(f)
    d = mj.decode(jsonstring, classes=rdrone.Drone)  # jsonpickle.decode(jsonstring, classes=Drone)
    assert type(d) is rdrone.Drone and type(d.controller) is rbrain.NeuralNetwork and type(d.sensors[0]) is rsensor.Sensor
    assert type(d.quadcopter).__name__ == "QuadcopterPhysics" and d.quadcopter.G == 0.4 * 9.81 and d.dh.visualize is False
    assert d.quadcopter.angle_control_to_thrust.tolist() == [[1, -1, -1], [-1, -1, 1], [-1, 1, -1], [1, 1, 1]]
    assert all(np.array_equal(a, b) and a.dtype == np.float64 for a, b in zip(d.controller.weights, ours.controller.weights))
    d.reset()
    with contextlib.redirect_stdout(io.StringIO()):
        reward, real_time, _ = ref_fly(d, idx=0)
    assert (reward, d.flight_time, d.crashed) == (-4077, 170, True) and abs(real_time - 1.69) < 1e-12
    # and the other way round through the same pickler: reference Drone -> jsonpickle text -> our loader
    with open(tmp_path / "ref.json", "w") as f:
        json.dump(mj.encode(d), f)
    back = checkpoint.load_drone(tmp_path / "ref.json")
    assert np.array_equal(back.controller.genome(), ours.controller.genome()) and back.reward == -4077
