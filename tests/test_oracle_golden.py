"""The CPU oracle (oracle/ld_oracle.c) against fixtures produced by the real reference.

Fixtures: tests/golden/*.npz, written by oracle/gen_golden.py from /root/reference
(SURVEY.md §8c known answers are asserted literally as well).  Integer outcomes must be
identical; continuous state must agree within STATE_RTOL/ATOL — the reference's BLAS/SIMD
transcendentals are not bit-reproducible across hosts (see ld_oracle.c header).
"""
import numpy as np
import pytest

from learning_drones_b200._abi import GENOME_LEN, make_params

from _parity import SET_STATE_FRAC, SETS, check_records, conditioning


@pytest.mark.parametrize("name,kw,min_robust", SETS)
@pytest.mark.parametrize("use_fma", [True, False])
def test_oracle_matches_reference(oracle, golden, name, kw, min_robust, use_fma):
    g = golden(name)
    p = make_params(int(g["timesteps"]), **kw)
    cond = conditioning(oracle, p, g["genomes"])
    rec, state = oracle.fly_mlp(p, g["genomes"], use_fma=use_fma, want_state=True)
    robust = check_records(rec, state, g, cond, SET_STATE_FRAC.get(name))
    assert robust.mean() >= min_robust


def test_known_answers_from_survey(golden):
    """SURVEY.md §8c literal values, so a regenerated fixture cannot drift silently."""
    g = golden("seed0_single")
    assert (int(g["reward"][0]), float(g["real_time"][0])) == (-4077, 1.69)
    assert int(g["flight_time"][0]) == 170 and bool(g["crashed"][0]) and int(g["coins"][0]) == 0
    assert float(g["distance"][0]) == pytest.approx(11.208182019123052, rel=1e-12)
    np.testing.assert_allclose(g["state"][0, :3], [0.79547323, 2.71876033, -0.13613557], atol=1e-8)
    g = golden("seed12345_gen0")
    assert g["reward"].tolist() == [-4501, -4736, -4164, -4585, -5294, -5346, -4028, -4294, -4431, -4218,
                                    -4038, -4187, -4581, -4186, -4503, -4140, -4504, -4279, -4191, -4302]
    assert g["flight_time"].tolist() == [168, 224, 160, 172, 190, 190, 149, 152, 154, 158,
                                         169, 173, 188, 197, 159, 205, 173, 154, 172, 202]
    g = golden("hover_2000")
    assert int(g["reward"][0]) == 3300 and not bool(g["crashed"][0]) and float(g["real_time"][0]) == 1999 * 0.01
    np.testing.assert_allclose(g["state"][0, :3], [0.0, 0.0, 8.03898111], atol=1e-8)


@pytest.mark.parametrize("name", ["seed0_single", "hover_2000"])
def test_oracle_trajectory(oracle, golden, name):
    g = golden(name)
    p = make_params(int(g["timesteps"]))
    rec, tr = oracle.trace_mlp(p, g["genomes"][0])
    ref = g["trace0"]
    n = ref.shape[0]
    assert n == int(g["flight_time"][0]) == int(rec["flight_time"])
    np.testing.assert_allclose(tr[:n], ref, rtol=1e-7, atol=1e-9)
    # free-fall fingerprint of the float32 force vector: a_z = -9.80999947
    if name == "seed0_single":
        assert tr[0, 14] == ref[0, 14] == pytest.approx(-9.80999947, abs=1e-8)
        assert tr[0, 14] != -9.81


def test_closed_form_reward(oracle, golden):
    """SURVEY §8a invariant: reward = 2*ft + coin values - 3000*crashed - unstable - 100*floor(dist)."""
    g = golden("selected_2500")
    p = make_params(2500)
    rec, _ = oracle.fly_mlp(p, g["genomes"])
    coins = rec["coins"].astype(np.int64)
    expect = (2 * rec["flight_time"].astype(np.int64) + 2000 * (coins >= 1) + 1000 * np.maximum(coins - 1, 0)
              - 3000 * (rec["status"] & 1) - rec["unstable"] - 100 * np.floor(rec["distance_to_coin"]).astype(np.int64))
    assert rec["reward"].tolist() == expect.tolist()
    assert coins.max() >= 2  # the fixture exercises the coin schedule


def test_population_generators_match_fixture(golden):
    """bench/tests build populations with oracle.gen_golden's formulas; the fixture rows are their prefix."""
    from oracle import gen_golden as G

    rng = np.random.default_rng(2026)
    gs = G.survivors(65536, rng)[:96]
    gr = G.random_brains(65536, rng)[:512]
    assert np.array_equal(gs, golden("survivors_2000")["genomes"])
    assert np.array_equal(gr, golden("random_2000")["genomes"])
    assert gs.shape[1] == GENOME_LEN


def test_oracle_thread_count_independent(oracle, golden):
    g = golden("random_2000")
    p = make_params(2000)
    a, sa = oracle.fly_mlp(p, g["genomes"][:64], nthreads=1, want_state=True)
    b, sb = oracle.fly_mlp(p, g["genomes"][:64], nthreads=4, want_state=True)
    assert a.tobytes() == b.tobytes() and sa.tobytes() == sb.tobytes()


def test_edge_cases(oracle):
    p = make_params(0)
    rec, st = oracle.fly_mlp(p, np.zeros((3, GENOME_LEN)), want_state=True)
    assert rec["flight_time"].tolist() == [0, 0, 0]
    # distance from (0,0,10) to (5,5,10) = 7.07 -> -700
    assert rec["reward"].tolist() == [-700] * 3
    rec, _ = oracle.fly_mlp(make_params(5), np.zeros((0, GENOME_LEN)))
    assert rec.shape == (0,)


def _large_fixture(golden):
    from oracle import gen_golden as G

    g = golden("random_2000_large")
    rng = np.random.default_rng(2026)
    G.survivors(65536, rng)
    lo, n = int(g["first_row"]), int(g["n"])
    genomes = G.random_brains(65536, rng)[lo:lo + n]
    assert float(genomes.sum()) == float(g["checksum"])  # the seed reproduces the brains the reference flew
    return g, genomes


def test_oracle_matches_reference_large(oracle, golden):
    """8,192 random drones flown by the real reference (results-only fixture)."""
    g, genomes = _large_fixture(golden)
    p = make_params(int(g["timesteps"]))
    cond = conditioning(oracle, p, genomes)
    rec, state = oracle.fly_mlp(p, genomes, want_state=True)
    robust = cond < 1e-9
    assert robust.mean() > 0.98
    for k in ("reward", "coins", "flight_time"):
        assert rec[k][robust].tolist() == g[k][robust].tolist()
    assert ((rec["status"] & 1) != 0)[robust].tolist() == g["crashed"][robust].tolist()
    dev = np.max(np.abs(state[robust, :3] - g["position"][robust]) / (1 + np.abs(g["position"][robust])), axis=1)
    assert np.median(dev) < 1e-10 and (dev < 1e-6).mean() > 0.99
    # among ALL drones (chaotic ones included) only a handful may differ
    same = (rec["reward"] == g["reward"]) & (rec["flight_time"] == g["flight_time"])
    assert same.mean() > 0.995


def test_oracle_continues_from_drone_state(oracle, golden):
    """fly() on Drones that are not in the reset state (constructor arguments; a second flight without reset()):
    the reference continues from the Drone's state and adds to its counters (flight/__init__.py:35-45)."""
    g = golden("start_states")
    n = len(g["genomes"])
    rows = np.zeros((n, 28))
    rows[:, 0:3], rows[:, 3:6], rows[:, 6:9], rows[:, 9:12] = g["init_pos"], g["init_vel"], g["init_angle"], g["init_angle_vel"]
    r1, s1 = oracle.fly_mlp(make_params(300), g["genomes"], want_state=True, init_state=rows)
    assert r1["reward"].tolist() == g["first_reward"].tolist() and r1["flight_time"].tolist() == g["first_flight_time"].tolist()
    assert r1["coins"].tolist() == g["first_coins"].tolist()
    np.testing.assert_allclose(s1, g["first_state"], rtol=1e-7, atol=1e-8)
    r2, s2 = oracle.fly_mlp(make_params(200), g["genomes"], want_state=True, init_state=s1)
    assert (r1["reward"] + r2["reward"]).tolist() == g["second_reward"].tolist()
    assert (r1["flight_time"] + r2["flight_time"]).tolist() == g["second_flight_time"].tolist()
    np.testing.assert_allclose((r2["flight_time"] - 1) * 0.01, g["second_real_time"], atol=1e-12)
    ok = np.isfinite(g["second_state"]).all(axis=1)
    np.testing.assert_allclose(s2[ok], g["second_state"][ok], rtol=1e-6, atol=1e-7)
