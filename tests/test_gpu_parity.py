"""GPU parity tests: the CUDA episode kernel (through the C-ABI) against the reference fixtures
and the CPU oracle.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

from _parity import COND_ROBUST, SET_STATE_FRAC, SETS, check_records, conditioning, state_dev  # noqa
from learning_drones_b200._abi import GENOME_LEN, RECORD_DTYPE, STATE_LEN, make_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pop():
    from learning_drones_b200 import population

    return population


@pytest.fixture(scope="module")
def gen():
    from oracle import gen_golden

    return gen_golden


@pytest.mark.parametrize("name,kw,min_robust", SETS)
def test_kernel_matches_reference_fixture(pop, oracle, golden, parity_report, name, kw, min_robust):
    g = golden(name)
    p = make_params(int(g["timesteps"]), **kw)
    cond = conditioning(oracle, p, g["genomes"])
    rec, state = pop.fly_population_host(g["genomes"], p, want_state=True)
    robust = check_records(rec, state, g, cond, SET_STATE_FRAC.get(name))
    assert robust.mean() >= min_robust
    differ = (rec["reward"] != g["reward"]) | (rec["flight_time"] != g["flight_time"]) | (rec["coins"] != g["coins"])
    parity_report(f"GPU vs REFERENCE fixture {name}: {int(differ.sum())} of {len(differ)} drones differ "
                  f"({int((differ & robust).sum())} of them well-conditioned; coin collectors differing: "
                  f"{int((differ & (g['coins'] > 0)).sum())} of {int((g['coins'] > 0).sum())})")


@pytest.mark.parametrize("family,n,H", [("survivors", 4096, 2000), ("random", 4096, 2000), ("random", 20000, 2500),
                                        ("survivors", 65536, 2000), ("random", 65536, 2000)])  # last two: config 3 at FULL size
def test_kernel_matches_oracle_subsample(pop, oracle, gen, parity_report, family, n, H):
    """parity on a subsample of the config-3 populations against the CPU oracle."""
    rng = np.random.default_rng(2026)
    s = gen.survivors(65536, rng)
    r = gen.random_brains(65536, rng)
    genomes = (s if family == "survivors" else r)[:n] if n <= 65536 else None
    p = make_params(H)
    cond = conditioning(oracle, p, genomes)
    ref, ref_state = oracle.fly_mlp(p, genomes, want_state=True)
    rec, state = pop.fly_population_host(genomes, p, want_state=True)
    g = dict(reward=ref["reward"], coins=ref["coins"], flight_time=ref["flight_time"], status=ref["status"],
             distance_to_coin=ref["distance_to_coin"], state=ref_state)
    robust = check_records(rec, state, g, cond)
    assert robust.mean() > 0.98
    # the ill-conditioned remainder may differ, but only rarely
    same = (rec["reward"] == ref["reward"]) & (rec["flight_time"] == ref["flight_time"])
    parity_report(f"GPU vs oracle, {family} {n} drones x {H} steps: {int((~same).sum())} of {n} drones differ "
                  f"({int((~same & robust).sum())} of them well-conditioned, cond < {COND_ROBUST:g}); rank order of the "
                  f"agreeing drones identical")
    assert same.mean() > 0.995


def test_trace_matches_oracle(pop, oracle, golden):
    for name in ("seed0_single", "hover_2000"):
        g = golden(name)
        p = make_params(int(g["timesteps"]))
        rec, tr = pop.trace_drone(g["genomes"][0], p)
        orec, otr = oracle.trace_mlp(p, g["genomes"][0])
        n = int(orec["flight_time"])
        assert int(rec["flight_time"]) == n == g["trace0"].shape[0]
        assert int(rec["reward"]) == int(orec["reward"]) == int(g["reward"][0])
        np.testing.assert_allclose(tr, otr[:n], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(tr, g["trace0"], rtol=1e-7, atol=1e-9)
        assert tr[0, 14] == otr[0, 14]  # float32 fingerprint of step 0 is bit-exact


def test_pipelined_host_submissions(pop, gen):
    """ld_fly_mlp_f64_host_submit / _wait: two submissions in flight per thread, results identical to the blocking call
    whatever the order of the waits; a third submission, a stale ticket and a blocking call in between are refused."""
    from learning_drones_b200._lib import LdError

    rng = np.random.default_rng(21)
    pops = [np.concatenate([gen.survivors(9000, rng), gen.random_brains(15000 + 1000 * i, rng)]) for i in range(5)]
    p = make_params(500)
    want = [pop.fly_population_host(g, p, want_state=True) for g in pops]
    a = pop.fly_population_host_submit(pops[0], p, want_state=True)
    b = pop.fly_population_host_submit(pops[1], p, want_state=True)
    with pytest.raises(LdError):
        pop.fly_population_host_submit(pops[2], p)       # pipeline full
    with pytest.raises(LdError):
        pop.fly_population_host(pops[2], p)              # the blocking call needs the pipeline drained
    rb, sb = b.result()                                  # waits may come in any order
    ra, sa = a.result()
    assert ra.tobytes() == want[0][0].tobytes() and rb.tobytes() == want[1][0].tobytes()
    assert np.array_equal(sa, want[0][1]) and np.array_equal(sb, want[1][1])
    assert a.result()[0] is ra                           # idempotent
    # steady-state pipeline: submit k+1, then collect k
    pending, got = None, []
    for g in pops:
        nxt = pop.fly_population_host_submit(g, p)
        if pending is not None:
            got.append(pending.result()[0])
        pending = nxt
    got.append(pending.result()[0])
    assert [r.tobytes() for r in got] == [w[0].tobytes() for w in want]
    assert pop.fly_population_host(pops[0], p)[0].tobytes() == want[0][0].tobytes()  # drained: blocking call works again
    import ctypes as C

    from learning_drones_b200 import _lib

    assert _lib.load().ld_fly_mlp_f64_host_wait(C.c_int64(12345 << 8)) != 0    # unknown ticket


def test_device_api_equals_host_api(pop, gen):
    import torch

    rng = np.random.default_rng(7)
    genomes = gen.random_brains(5000, rng)
    p = make_params(600)
    rec_h, st_h = pop.fly_population_host(genomes, p, want_state=True)
    gt = torch.from_numpy(genomes).cuda()
    rec_d, st_d = pop.fly_population_device(gt, p, want_state=True)
    assert pop.records_to_numpy(rec_d).tobytes() == rec_h.tobytes()
    assert st_d.cpu().numpy().tobytes() == st_h.tobytes()
    # deterministic: a second launch (different lane/queue interleaving) gives identical bytes
    rec_d2, _ = pop.fly_population_device(gt, p)
    assert torch.equal(rec_d, rec_d2)
    # permutation invariance: drones are independent
    perm = rng.permutation(len(genomes))
    rec_p, _ = pop.fly_population_host(genomes[perm], p)
    assert rec_p.tobytes() == rec_h[perm].tobytes()


@pytest.mark.parametrize("slice_steps", [0, 64, 150, 599, 4000, -40, -1000000])  # the last one = LD_SLICE_STAGGER
def test_episode_segments_do_not_change_results(pop, gen, slice_steps):
    """ld_params.slice_steps is scheduling only: with more drones than lanes, drones are parked at segment
    boundaries and resumed on other lanes — records and final states stay bit-identical to the unsegmented run."""
    import torch

    from learning_drones_b200 import _lib

    wave = int(_lib.load().ld_wave_size())
    rng = np.random.default_rng(99)
    n = 2 * wave + 4321
    genomes = np.empty((n, GENOME_LEN))
    genomes[0::2] = gen.survivors((n + 1) // 2, rng)      # fly all 600 steps
    genomes[1::2] = gen.random_brains(n // 2, rng)         # mostly crash within a segment or two
    gt = torch.from_numpy(genomes).cuda()
    p = make_params(600)
    p.slice_steps = -1
    rec0, st0 = pop.fly_population_device(gt, p, want_state=True)
    r0 = pop.records_to_numpy(rec0)
    assert (r0["flight_time"][0::2] == 600).all() and (r0["flight_time"][1::2] < 600).any()
    p.slice_steps = slice_steps
    for _ in range(2):  # twice: the interleaving of lanes, parks and resumes differs from launch to launch
        rec1, st1 = pop.fly_population_device(gt, p, want_state=True)
        assert torch.equal(rec0, rec1)
        assert st0.cpu().numpy().tobytes() == st1.cpu().numpy().tobytes()
    rec2, _ = pop.fly_population_device(gt, p)  # the kernel variant without the state output
    assert torch.equal(rec0, rec2)
    # host buffers: one launch that starts before its rows have arrived (pageable and pinned source)
    rec_h, st_h = pop.fly_population_host(genomes, p, want_state=True)
    assert rec_h.tobytes() == r0.tobytes() and st_h.tobytes() == st0.cpu().numpy().tobytes()
    pinned = torch.from_numpy(genomes).pin_memory()
    rec_p, _ = pop.fly_population_host(pinned.numpy(), p)
    assert rec_p.tobytes() == r0.tobytes()


@pytest.mark.parametrize("extra,slice_steps,batch", [(1, 7, 1), (33, 13, 32), (5000, 1, 0), (77, -3, 5)])
def test_episode_segments_stress(pop, gen, extra, slice_steps, batch):
    """Tiny segments (down to one step), a population just above one wave, extreme refill batches."""
    import torch

    from learning_drones_b200 import _lib

    wave = int(_lib.load().ld_wave_size())
    rng = np.random.default_rng(extra)
    n = wave + extra
    genomes = np.concatenate([gen.random_brains(n - n // 3, rng), gen.survivors(n // 3, rng)])[rng.permutation(n)]
    gt = torch.from_numpy(genomes).cuda()
    p = make_params(60)
    p.slice_steps = -1
    rec0, st0 = pop.fly_population_device(gt, p, want_state=True)
    p.slice_steps, p.refill_batch = slice_steps, batch
    rec1, st1 = pop.fly_population_device(gt, p, want_state=True)
    assert torch.equal(rec0, rec1) and st0.cpu().numpy().tobytes() == st1.cpu().numpy().tobytes()
    rec_h, _ = pop.fly_population_host(genomes, p)
    assert rec_h.tobytes() == pop.records_to_numpy(rec0).tobytes()


@pytest.mark.parametrize("n", [0, 1, 31, 33, 95, 97, 14207, 14209])
def test_ragged_population_sizes(pop, oracle, gen, n):
    rng = np.random.default_rng(100 + n)
    genomes = gen.random_brains(max(n, 1), rng)[:n]
    p = make_params(400)
    rec, st = pop.fly_population_host(genomes, p, want_state=True)
    assert rec.shape == (n,) and st.shape == (n, STATE_LEN)
    if n:
        ref, _ = oracle.fly_mlp(p, genomes)
        cond = conditioning(oracle, p, genomes)
        ok = cond < COND_ROBUST
        assert rec["reward"][ok].tolist() == ref["reward"][ok].tolist()
        assert rec["flight_time"][ok].tolist() == ref["flight_time"][ok].tolist()


def test_zero_timesteps_and_flags(pop, oracle, gen):
    rng = np.random.default_rng(3)
    genomes = gen.random_brains(200, rng)
    for kw in (dict(timesteps=0), dict(timesteps=1), dict(timesteps=300, reward_time=False),
               dict(timesteps=300, punish_time=True), dict(timesteps=300, punish_time=True, reward_time=False)):
        p = make_params(**kw)
        rec, st = pop.fly_population_host(genomes, p, want_state=True)
        ref, rst = oracle.fly_mlp(p, genomes, want_state=True)
        cond = conditioning(oracle, p, genomes)
        ok = cond < COND_ROBUST
        assert rec["reward"][ok].tolist() == ref["reward"][ok].tolist(), kw
        assert rec["flight_time"][ok].tolist() == ref["flight_time"][ok].tolist(), kw
        assert state_dev(st[ok], rst[ok]).max() < 1e-6


def test_wind_and_payload(pop, oracle, gen):
    rng = np.random.default_rng(5)
    genomes = gen.random_brains(512, rng)
    p = make_params(500, wind_speed=[0.5, -2.0, 0.25], mass_payload=0.05, x_payload=0.01, y_payload=-0.02)
    rec, st = pop.fly_population_host(genomes, p, want_state=True)
    ref, rst = oracle.fly_mlp(p, genomes, want_state=True)
    cond = conditioning(oracle, p, genomes)
    g = dict(reward=ref["reward"], coins=ref["coins"], flight_time=ref["flight_time"], status=ref["status"],
             distance_to_coin=ref["distance_to_coin"], state=rst)
    assert check_records(rec, st, g, cond).mean() > 0.95
    # and it is not the no-wind result
    rec0, _ = pop.fly_population_host(genomes, make_params(500))
    assert (rec0["reward"] != rec["reward"]).any()


def test_reduce_records(pop, gen):
    import torch

    rng = np.random.default_rng(9)
    genomes = torch.from_numpy(gen.random_brains(30000, rng)).cuda()
    rec, _ = pop.fly_population_device(genomes, make_params(500))
    stats = pop.reduce_records(rec)
    r = pop.records_to_numpy(rec)
    assert stats["active_steps"] == int(r["flight_time"].sum())
    assert stats["best_reward"] == int(r["reward"].max())
    assert stats["best_index"] == int(np.argmax(r["reward"]))
    assert stats["coins"] == int(r["coins"].sum())
    assert stats["crashed"] == int((r["status"] & 1).sum())


def test_full_size_properties(pop, gen):
    """Config 3 at full size (65,536 x 2,000): size-independent invariants."""
    import torch

    rng = np.random.default_rng(2026)
    s = gen.survivors(65536, rng)
    p = make_params(2000)
    gt = torch.from_numpy(s).cuda()
    rec, _ = pop.fly_population_device(gt, p)
    r = pop.records_to_numpy(rec)
    assert (r["flight_time"] == 2000).all() and (r["status"] == 0).all()  # the (S) family survives
    coins = r["coins"].astype(np.int64)
    expect = (2 * r["flight_time"].astype(np.int64) + 2000 * (coins >= 1) + 1000 * np.maximum(coins - 1, 0)
              - 3000 * (r["status"] & 1) - r["unstable"] - 100 * np.floor(r["distance_to_coin"]).astype(np.int64))
    assert r["reward"].tolist() == expect.tolist()
    assert len(np.unique(r["reward"])) > 20  # non-trivial fitness ranking
    stats = pop.reduce_records(rec)
    assert stats["active_steps"] == 65536 * 2000


def test_fly_drop_in(golden):
    """Reference-API call: fly(drone) on Drone objects (flight/__init__.py contract)."""
    from learning_drones_b200.drone.drone import Drone
    from learning_drones_b200.flight import fly, fly_population

    g = golden("seed0_single")
    np.random.seed(0)
    d = Drone(visualize=False, n_servers=0)
    # same global-RNG draws as the reference => same brain as the fixture
    assert np.array_equal(d.controller.genome(), g["genomes"][0])
    reward, real_time, idx = fly(d, idx=0)
    assert (reward, real_time, idx) == (-4077, 1.69, 0)
    assert isinstance(reward, int) and d.crashed and d.coins == 0 and d.flight_time == 170
    assert d.distance_to_coin == pytest.approx(11.208182019123052, rel=1e-9)
    np.testing.assert_allclose(d.position.ravel(), g["state"][0, :3], rtol=1e-7)
    assert d.position.shape == (3, 1) and d.thrust.shape == (4, 1)
    # a second fly() without reset() continues from the crashed state (z <= 0: crashes again at once) and adds to
    # the counters, like the reference does (flight/__init__.py:35-45); evolution.py resets between generations
    reward2, real_time2, _ = fly(d)
    assert d.flight_time == 171 and reward2 < reward - 3000 and real_time2 == 0.0 and d.crashed
    d.reset()
    assert fly(d, idx=5)[0] == -4077

    g = golden("seed12345_gen0")
    np.random.seed(12345)
    drones = [Drone(visualize=False, n_servers=0) for _ in range(20)]
    res = fly_population(drones)
    assert [r[0] for r in res] == g["reward"].tolist()
    assert [d.flight_time for d in drones] == g["flight_time"].tolist()
    np.testing.assert_allclose([r[1] for r in res], g["real_time"])


def test_fly_error_behaviour():
    """Exceptions of the reference: int(nan) -> ValueError (drone/drone.py:14), timesteps=0 ->
    UnboundLocalError on `real_time` (flight/__init__.py:99)."""
    from learning_drones_b200.drone.drone import Drone
    from learning_drones_b200.flight import fly

    np.random.seed(4)
    d = Drone(visualize=False, n_servers=0)
    g = d.controller.genome()
    g[3] = np.nan  # a NaN weight poisons the thrust, then the whole state
    d.controller.set_genome(g)
    with pytest.raises(ValueError):
        fly(d, timesteps=50)
    d2 = Drone(visualize=False, n_servers=0)
    with pytest.raises(UnboundLocalError):
        fly(d2, timesteps=0)
    with pytest.raises(NotImplementedError):
        fly(Drone(visualize=False, n_servers=0), visualize=True)
    d3 = Drone(visualize=False, n_servers=0)
    d3.controller.weights = d3.controller.weights[:2]  # wrong brain shape
    with pytest.raises(ValueError):
        fly(d3)


def test_config4_size_properties(pop, gen):
    """Config 4 population size (1,048,576 drones) on one GPU: invariants that need no oracle."""
    import torch

    rng = np.random.default_rng(4)
    base = gen.random_brains(8192, rng)
    big = np.tile(base, (128, 1))  # 1 M drones = 128 copies of 8192 brains
    p = make_params(400)
    gt = torch.from_numpy(big).cuda()
    rec, _ = pop.fly_population_device(gt, p)
    r = pop.records_to_numpy(rec).reshape(128, 8192)
    # every copy of a brain gets the identical record, wherever the queue placed it
    assert all(r[k].tobytes() == r[0].tobytes() for k in range(1, 128))
    coins = r[0]["coins"].astype(np.int64)
    expect = (2 * r[0]["flight_time"].astype(np.int64) + 2000 * (coins >= 1) + 1000 * np.maximum(coins - 1, 0)
              - 3000 * (r[0]["status"] & 1) - r[0]["unstable"] - 100 * np.floor(r[0]["distance_to_coin"]).astype(np.int64))
    assert r[0]["reward"].tolist() == expect.tolist()
    stats = pop.reduce_records(rec)
    assert stats["active_steps"] == 128 * int(r[0]["flight_time"].sum())
    assert stats["best_reward"] == int(r[0]["reward"].max()) and stats["best_index"] == int(np.argmax(r[0]["reward"]))


def test_concurrent_callers(pop, gen):
    """Host threads and CUDA streams calling the library at the same time: every kernel owns its SMs'
    Tensor Memory exclusively, so concurrent launches must serialise per SM without deadlock and
    give the results of a lone call."""
    import threading

    import torch

    rng = np.random.default_rng(21)
    genomes = [gen.random_brains(3000 + 500 * i, rng) for i in range(4)]
    p = make_params(500)
    alone = [pop.fly_population_host(g, p)[0] for g in genomes]
    out = [None] * 4

    def work(i):
        out[i] = pop.fly_population_host(genomes[i], p)[0]

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in threads]
    [t.join(timeout=120) for t in threads]
    assert all(not t.is_alive() for t in threads)
    assert all(out[i].tobytes() == alone[i].tobytes() for i in range(4))
    # device API on several torch streams at once
    gts = [torch.from_numpy(g).cuda() for g in genomes]
    streams = [torch.cuda.Stream() for _ in range(4)]
    recs = []
    for g, s in zip(gts, streams):
        with torch.cuda.stream(s):
            recs.append(pop.fly_population_device(g, p)[0])
    torch.cuda.synchronize()
    assert all(pop.records_to_numpy(r).tobytes() == a.tobytes() for r, a in zip(recs, alone))


def test_large_angle_fallback_path(pop, oracle, gen):
    """Angles are never wrapped: beyond |x| = 105615 the kernel's fast range reduction hands
    over to CUDA's libm (warp-uniform fallback branch).  Spin the drones up so that the branch is taken
    and compare with the oracle (glibc) — a regular, non-chaotic free fall, so tolerances stay tight."""
    genomes = np.zeros((64, GENOME_LEN))
    genomes[:, 270 + 10::11] = -50.0  # output biases: all four thrusts ~ 0 => no torque, pure spin + free fall
    for rate in ([3e6, 0.0, 0.0], [0.0, 0.0, -5e6], [2e6, 1e3, 4e6]):
        p = make_params(100, initial_angle_vel=rate)
        rec, st = pop.fly_population_host(genomes, p, want_state=True)
        ref, rst = oracle.fly_mlp(p, genomes, want_state=True)
        assert np.abs(rst[:, 6:9]).max() > 1.5e5            # the fallback range was reached
        assert rec["flight_time"].tolist() == ref["flight_time"].tolist()
        assert rec["reward"].tolist() == ref["reward"].tolist()
        np.testing.assert_allclose(st[:, :12], rst[:, :12], rtol=1e-9, atol=1e-9)


def test_saturating_brains_and_nonfinite_status(pop, oracle, gen):
    """Huge weights drive exp(-0.4 x) into overflow / underflow: thrust must be exactly 0 or 1 like the
    reference (1/(inf+1) = 0).  NaN weights must surface as LD_STATUS_NONFINITE."""
    from learning_drones_b200._abi import STATUS_NONFINITE

    rng = np.random.default_rng(23)
    genomes = gen.random_brains(2000, rng) * rng.choice([1.0, 30.0, 1e3, 1e5], size=(2000, 1))
    p = make_params(400)
    rec, st = pop.fly_population_host(genomes, p, want_state=True)
    ref, rst = oracle.fly_mlp(p, genomes, want_state=True)
    cond = conditioning(oracle, p, genomes)
    ok = cond < COND_ROBUST
    assert ok.mean() > 0.9
    assert rec["reward"][ok].tolist() == ref["reward"][ok].tolist()
    assert rec["flight_time"][ok].tolist() == ref["flight_time"][ok].tolist()
    sat = np.isin(rst[:, 24:28], (0.0, 1.0)).all(axis=1) & ok
    assert sat.sum() > 100 and np.array_equal(st[sat, 24:28], rst[sat, 24:28])  # exact 0 / 1 thrusts
    bad = genomes[:8].copy()
    bad[:, 5] = np.nan
    rec, _ = pop.fly_population_host(bad, make_params(50))
    ref, _ = oracle.fly_mlp(make_params(50), bad)
    assert ((rec["status"] & STATUS_NONFINITE) != 0).all() and ((ref["status"] & STATUS_NONFINITE) != 0).all()
    assert rec["flight_time"].tolist() == ref["flight_time"].tolist() == [50] * 8  # NaN never crashes (drone.py:150)


def test_kernel_matches_reference_large(pop, oracle, golden, parity_report):
    """The CUDA kernel against 8,192 random drones flown by the REAL reference (results-only fixture)."""
    from test_oracle_golden import _large_fixture

    g, genomes = _large_fixture(golden)
    p = make_params(int(g["timesteps"]))
    cond = conditioning(oracle, p, genomes)
    rec, state = pop.fly_population_host(genomes, p, want_state=True)
    robust = cond < COND_ROBUST
    for k in ("reward", "coins", "flight_time"):
        assert rec[k][robust].tolist() == g[k][robust].tolist()
    assert ((rec["status"] & 1) != 0)[robust].tolist() == g["crashed"][robust].tolist()
    dev = np.max(np.abs(state[robust, :3] - g["position"][robust]) / (1 + np.abs(g["position"][robust])), axis=1)
    assert np.median(dev) < 1e-10 and (dev < 1e-6).mean() > 0.99
    same = (rec["reward"] == g["reward"]) & (rec["flight_time"] == g["flight_time"])
    parity_report(f"GPU vs REFERENCE, 8192 random drones x {int(g['timesteps'])} steps flown by the real reference: {int((~same).sum())} of "
                  f"{len(same)} drones differ ({int((~same & robust).sum())} of them well-conditioned; {int(robust.sum())} well-conditioned in all)")
    assert same.mean() > 0.995
    # fitness rank order of the whole generation (north star): identical on the robust drones
    a, b = rec["reward"][robust], g["reward"][robust]
    assert np.argsort(-a, kind="stable").tolist() == np.argsort(-b, kind="stable").tolist()


def test_whole_evolution_loop_drop_in(golden):
    """BASELINE config 2: the generation loop of evolve/evolution.py (seed 12345, 20 drones, 6 generations)
    with this package's Drone / fly_population / evolve.util in place of the reference's — scores, flight
    times, best index and global best of EVERY generation must equal what the reference produced
    (fixture evolution_seed12345.npz: the same loop driven through the reference's own classes)."""
    from oracle.gen_golden import evolution_loop

    from learning_drones_b200.drone.drone import Drone
    from learning_drones_b200.evolve import util
    from learning_drones_b200.flight import fly_population

    g = golden("evolution_seed12345")
    hist = evolution_loop(Drone, fly_population, util, seed=12345, n_generations=len(g["scores"]))
    for gen, h in enumerate(hist):
        assert h["scores"] == g["scores"][gen].tolist(), gen
        np.testing.assert_allclose(h["times"], g["times"][gen], rtol=0, atol=1e-12)
        assert h["best_index"] == int(g["best_index"][gen]) and h["global_best"] == int(g["global_best"][gen])


def test_stepping_a_drone_by_hand_equals_fly(golden):
    """Drone.measure() + Drone.translate_input_to_thrust() driven by the reference's own loop
    (flight/__init__.py:39-93, restated here) reproduce fly() bit for bit and hit the reference's numbers."""
    from learning_drones_b200.drone.drone import REWARD_COIN_DISTANCE, REWARD_TIME_PASSED, Drone
    from learning_drones_b200.drone.quadcopter.parameters import initial_wind_speed
    from learning_drones_b200.flight import fly
    from learning_drones_b200.objects.coins import Coin

    g = golden("seed0_single")
    np.random.seed(0)
    a, b = Drone(visualize=False, n_servers=0), Drone(visualize=False, n_servers=0)
    b.controller.set_genome(a.controller.genome())
    coin = Coin(np.array([[5.0], [5.0], [10.0]]), 2, 2000)
    for time in range(2500):
        a.reward += REWARD_TIME_PASSED
        a.flight_time += 1
        if coin.is_in_vicinity(a.position):
            a.reward += coin.value
            a.coins += 1
            coin = Coin(coin.position + np.array([[5.0], [2.5], [0.0]]), 1, value=1000)
        if a.measure(wind_speed=initial_wind_speed):
            break
        a.translate_input_to_thrust(coin.position)
    a.reward += REWARD_COIN_DISTANCE(coin.distance_drone_to_coin(a.position))
    reward, real_time, _ = fly(b, idx=0)
    assert (a.reward, a.flight_time, a.crashed, a.coins) == (reward, b.flight_time, b.crashed, b.coins) == (-4077, 170, True, 0)
    for x, y in ((a.position, b.position), (a.velocity, b.velocity), (a.angle, b.angle), (a.angle_vel, b.angle_vel), (a.thrust, b.thrust)):
        assert np.array_equal(x, y)  # same device functions => bit-identical
    assert [s.current_acceleration for s in a.sensors] == [s.current_acceleration for s in b.sensors]
    assert [s.last_acceleration for s in a.sensors] == [s.last_acceleration for s in b.sensors]
    np.testing.assert_allclose(a.position.ravel(), g["state"][0, :3], rtol=1e-7)


def test_brain_forward_matches_reference_fixture(golden):
    """NeuralNetwork.feed_forward / translate_input_to_thrust / relu / sigmoid run in libldrones.so for ANY layer_spec
    (no host fallback); fixture = the reference's own NeuralNetwork incl. its default {1: 8, 2: 10, 3: 4}.
    Tolerance 1e-13 relative: the reference's np.dot is an OpenBLAS dgemv whose summation order is host specific."""
    from learning_drones_b200.evolve.brain import NeuralNetwork

    g = golden("brain_forward")
    for name in ("default", "drone", "small", "deep"):
        sizes = g[f"{name}_sizes"].tolist()
        spec = {i + 1: s for i, s in enumerate(sizes[1:])}
        nn = NeuralNetwork(n_inputs=sizes[0]) if name == "default" else NeuralNetwork(layer_spec=spec, n_inputs=sizes[0])
        assert [w.shape for w in nn.weights] == [(o, i + 1) for i, o in zip(sizes[:-1], sizes[1:])]
        o, ws = 0, []
        for i, n_out in zip(sizes[:-1], sizes[1:]):
            ws.append(g[f"{name}_weights"][o:o + n_out * (i + 1)].reshape(n_out, i + 1).copy())
            o += n_out * (i + 1)
        nn.weights = ws
        for x, y in zip(g[f"{name}_inputs"], g[f"{name}_outputs"]):
            got = nn.feed_forward(x)
            assert got.shape == y.shape
            np.testing.assert_allclose(got, y, rtol=1e-13, atol=1e-300)
    nn = NeuralNetwork()
    assert np.array_equal(nn.relu(g["act_x"]), g["act_relu"]) and np.array_equal(nn.relu(g["act_x"], thresh=2.5, alpha=0.5), g["act_relu_t"])
    np.testing.assert_allclose(nn.sigmoid(g["act_x"]), g["act_sigmoid"], rtol=4e-16 * 4)
    assert nn._relu(-1.0) == 0.0 and nn._relu(3.0) == 3.0
    drone_nn = NeuralNetwork(layer_spec={1: 10, 2: 10, 3: 4})
    drone_nn.set_genome(g["thrust_weights"])
    thrust = drone_nn.translate_input_to_thrust(*g["thrust_args"])
    np.testing.assert_allclose(thrust, g["thrust_out"], rtol=1e-13)
    with pytest.raises(ValueError):
        drone_nn.feed_forward(np.zeros((7, 1)))


def test_fly_continues_from_drone_state(pop, oracle, golden):
    """Drone(initial_pos=...) flies from its constructor arguments, and fly() on a Drone that already flew continues
    from its state and adds to its counters — like the reference (fixture: oracle/gen_golden.py --starts-only)."""
    from learning_drones_b200.drone.drone import Drone
    from learning_drones_b200.flight import fly, fly_population

    g = golden("start_states")
    n = len(g["genomes"])
    col = lambda a: np.asarray(a, dtype=np.float64).reshape(3, 1)
    drones = []
    for i in range(n):
        d = Drone(visualize=False, n_servers=0, initial_pos=col(g["init_pos"][i]), initial_vel=col(g["init_vel"][i]),
                  initial_angle=col(g["init_angle"][i]), initial_angle_vel=col(g["init_angle_vel"][i]))
        d.controller.set_genome(g["genomes"][i])
        assert np.array_equal(d.position, col(g["init_pos"][i])) and d.sensors[3].previous_pos == g["init_angle"][i][0]
        drones.append(d)
    out = fly_population(drones, timesteps=300)
    assert [o[0] for o in out] == g["first_reward"].tolist() == [d.reward for d in drones]
    assert [d.flight_time for d in drones] == g["first_flight_time"].tolist() and [d.crashed for d in drones] == g["first_crashed"].tolist()
    np.testing.assert_allclose([o[1] for o in out], g["first_real_time"], atol=1e-12)
    np.testing.assert_allclose(np.array([d._state_row() for d in drones]), g["first_state"], rtol=1e-7, atol=1e-8)
    # second flight WITHOUT reset: half through fly() one by one, half as a population
    second = [fly(d, timesteps=200, idx=i) for i, d in enumerate(drones[: n // 2])]
    second += [(r, t, n // 2 + i) for i, (r, t, _) in enumerate(fly_population(drones[n // 2:], timesteps=200))]
    assert [o[0] for o in second] == g["second_reward"].tolist() == [d.reward for d in drones]
    assert [d.flight_time for d in drones] == g["second_flight_time"].tolist() and [d.coins for d in drones] == g["second_coins"].tolist()
    assert [d.crashed for d in drones] == g["second_crashed"].tolist()
    np.testing.assert_allclose([o[1] for o in second], g["second_real_time"], atol=1e-12)
    np.testing.assert_allclose([d.distance_to_coin for d in drones], g["second_distance"], rtol=1e-6)
    # reset() goes back to the module-level initial state whatever the constructor was given (drone.py:177-218)
    drones[0].reset()
    assert drones[0].position.ravel().tolist() == [0.0, 0.0, 10.0] and drones[0].reward == 0
    # device entry point with per-drone start rows == oracle, mixed population larger than one wave of lanes
    import torch

    rng = np.random.default_rng(5)
    m = 20000
    genomes = np.concatenate([g["genomes"]] * (m // n + 1))[:m]
    rows = np.zeros((m, STATE_LEN))
    rows[:, 0:3] = rng.uniform(-3, 3, size=(m, 3)) + [0, 0, 20]
    rows[:, 3:12] = rng.uniform(-0.3, 0.3, size=(m, 9))
    rows[:, 12:18] = rng.uniform(-1, 1, size=(m, 6))
    rows[:, 24:28] = rng.uniform(0, 0.4, size=(m, 4))
    p = make_params(400)
    p.slice_steps = 100
    ref, rst = oracle.fly_mlp(p, genomes, want_state=True, init_state=rows)
    rec_t, st_t = pop.fly_population_device(torch.from_numpy(genomes).cuda(), p, want_state=True, init_state=torch.from_numpy(rows).cuda())
    rec = pop.records_to_numpy(rec_t)
    same = (rec["reward"] == ref["reward"]) & (rec["flight_time"] == ref["flight_time"])
    assert same.mean() > 0.995, same.mean()
    rec_h, _ = pop.fly_population_host(genomes, p, init_state=rows)
    assert rec_h.tobytes() == rec.tobytes()
