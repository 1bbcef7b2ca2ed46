"""PID mode (flight/collect_coins.py, config 1): CPU oracle against fixtures produced by driving the
reference's own PID / PIDControlUNnit / QuadcopterPhysics / Sensor / Coin classes
(oracle/gen_golden.py: ref_fly_pid), cross-checked there against the verbatim script's printed score."""
import numpy as np
import pytest

from learning_drones_b200._abi import STATUS_FINISHED, make_params, make_pid_params, paramset_to_pid_gains


def test_script_known_answer(golden):
    g = golden("pid_collect_coins")
    i11 = list(g["paramsets"]).index("11")
    assert int(g["script_final_score_11"]) == 5923           # printed by the real script
    assert int(g["score"][i11]) == 5923 and int(g["steps"][i11]) == 1577
    assert g["pick_steps"][i11].tolist() == [145, 635, 949, 1576]


def test_gain_wiring():
    g18 = np.arange(18, dtype=float)
    w = paramset_to_pid_gains(g18)
    assert w.shape == (6, 3)
    assert w[0].tolist() == [0, 3, 6] and w[1].tolist() == [1, 4, 7]
    assert w[2].tolist() == [2, 4, 8]  # z controller uses ki_y (collect_coins.py:181)
    assert w[3].tolist() == [9, 12, 15] and w[5].tolist() == [11, 14, 17]
    assert paramset_to_pid_gains(np.zeros((5, 18))).shape == (5, 6, 3)


@pytest.mark.parametrize("use_fma", [True, False])
def test_pid_oracle_matches_reference(oracle, golden, use_fma):
    g = golden("pid_collect_coins")
    p, q = make_params(), make_pid_params(int(g["timesteps"]))
    rec, state, _ = oracle.fly_pid(p, q, paramset_to_pid_gains(g["gains"]), use_fma=use_fma, want_state=True)
    assert rec["reward"].tolist() == g["score"].tolist()
    assert rec["flight_time"].tolist() == g["steps"].tolist()
    assert rec["coins"].tolist() == g["coins"].tolist()
    assert ((rec["status"] & STATUS_FINISHED) != 0).tolist() == g["finished"].tolist()
    dev = np.max(np.abs(state[:, :12] - g["state"][:, :12]) / (1 + np.abs(g["state"][:, :12])), axis=1)
    assert np.median(dev) < 1e-10 and dev.max() < 1e-6  # a stabilised flight is not chaotic
    np.testing.assert_allclose(state[:, 24:], g["state"][:, 24:], atol=1e-9)


def test_pid_oracle_trajectory(oracle, golden):
    g = golden("pid_collect_coins")
    i11 = list(g["paramsets"]).index("11")
    p, q = make_params(), make_pid_params(int(g["timesteps"]))
    rec, _, tr = oracle.fly_pid(p, q, paramset_to_pid_gains(g["gains"][i11 : i11 + 1]), want_trace=True)
    ref = g["trace11"]
    n = ref.shape[0]
    assert n == int(rec["flight_time"][0]) - 1  # the exit step breaks before the physics
    np.testing.assert_allclose(tr[:n], ref, rtol=1e-7, atol=1e-9)
    assert np.isnan(tr[n:]).all()


def test_paramset_table_matches_reference_fixture(golden):
    """drone/quadcopter/choose_paramset.py mirror: every named gain set equals what the reference's own get_paramset +
    convert_paramset_2_float returned when the fixture was generated (oracle/gen_golden.py `pid_gains`)."""
    from learning_drones_b200.drone.quadcopter import choose_paramset as cp
    from learning_drones_b200.population import paramset_gains

    g = golden("pid_collect_coins")
    names = [str(n) for n in g["paramsets"]]
    assert set(names) == set(cp.PARAMSET_NAMES) and len(names) == 10
    for i, name in enumerate(names):
        d = cp.get_paramset(name)
        assert len(d) == 18 and np.array_equal(np.array(cp.convert_paramset_2_float(d)), g["gains"][i])
    assert np.array_equal(paramset_gains(names), paramset_to_pid_gains(g["gains"][:10]))
    assert paramset_gains("11").shape == (1, 6, 3) and paramset_gains("11")[0, 2, 1] == 0.1  # z PID uses ki_y (collect_coins.py:181)
    with pytest.raises(Exception, match="Invalid Parameter Set"):
        cp.get_paramset("default")
