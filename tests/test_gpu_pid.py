"""GPU parity for the PID controller mode (flight/collect_coins.py batch) — kernel vs reference
fixtures and vs the CPU oracle on a larger batch of gain sets."""
import numpy as np
import pytest

from learning_drones_b200._abi import STATUS_FINISHED, make_params, make_pid_params, paramset_to_pid_gains

pytestmark = pytest.mark.gpu


def test_pid_kernel_matches_reference_fixture(golden):
    from learning_drones_b200 import population as pop

    g = golden("pid_collect_coins")
    p, q = make_params(), make_pid_params(int(g["timesteps"]))
    rec, state = pop.fly_pid_population_host(paramset_to_pid_gains(g["gains"]), p, q, want_state=True)
    i11 = list(g["paramsets"]).index("11")
    assert int(rec["reward"][i11]) == 5923 and int(rec["flight_time"][i11]) == 1577
    assert rec["reward"].tolist() == g["score"].tolist()
    assert rec["flight_time"].tolist() == g["steps"].tolist()
    assert rec["coins"].tolist() == g["coins"].tolist()
    assert ((rec["status"] & STATUS_FINISHED) != 0).tolist() == g["finished"].tolist()
    dev = np.max(np.abs(state[:, :12] - g["state"][:, :12]) / (1 + np.abs(g["state"][:, :12])), axis=1)
    assert np.median(dev) < 1e-10 and dev.max() < 1e-6
    np.testing.assert_allclose(state[:, 24:], g["state"][:, 24:], atol=1e-9)


def test_pid_kernel_matches_oracle_batch(oracle, golden):
    import torch

    from learning_drones_b200 import population as pop

    g = golden("pid_collect_coins")
    base = g["gains"][list(g["paramsets"]).index("11")]
    rng = np.random.default_rng(42)
    gains18 = base * rng.uniform(0.5, 1.6, size=(3000, 18))
    pg = paramset_to_pid_gains(gains18)
    p, q = make_params(), make_pid_params(3500)
    ref, rst, _ = oracle.fly_pid(p, q, pg, want_state=True)
    rec_t, st_t = pop.fly_pid_population_device(torch.from_numpy(pg).cuda(), p, q, want_state=True)
    rec, st = pop.records_to_numpy(rec_t), st_t.cpu().numpy()
    same = (rec["reward"] == ref["reward"]) & (rec["flight_time"] == ref["flight_time"]) & (rec["coins"] == ref["coins"])
    # The PID law is discontinuous (integral flush on a sign change, |e| < 2 window, clamps), so a
    # 1-ulp difference can flip a decision.  Same contract as the MLP mode (tests/_parity.py): gain
    # sets whose outcome is insensitive to the oracle's own fused/unfused variants must agree exactly.
    ref0, rst0, _ = oracle.fly_pid(p, q, pg, use_fma=False, want_state=True)
    cond = np.max(np.abs(rst - rst0)[:, :12] / (1 + np.abs(rst0[:, :12])), axis=1)
    ok = (cond < 1e-6) & (ref0["reward"] == ref["reward"])
    assert ok.sum() > 2000 and same[ok].all(), (ok.sum(), (~same[ok]).sum())
    assert same.mean() > 0.97
    dev = np.max(np.abs(st[ok, :12] - rst[ok, :12]) / (1 + np.abs(rst[ok, :12])), axis=1)
    assert np.median(dev) < 1e-8 and (dev < 1e-3).mean() > 0.99
    # host entry point == device entry point
    rec_h, _ = pop.fly_pid_population_host(pg, p, q)
    assert rec_h.tobytes() == rec.tobytes()
    # ragged / empty
    assert pop.fly_pid_population_host(pg[:0], p, q)[0].shape == (0,)
    assert pop.fly_pid_population_host(pg[:33], p, q)[0].tobytes() == rec[:33].tobytes()


def test_config1_by_paramset_name():
    """BASELINE.json configs[0]: `run_simulation_without_viz.sh 11` — the headless coin-collection flight with the
    reference's gain set "11", named through the API (no fixture involved): Final Score 5923 after 1577 steps."""
    from learning_drones_b200 import population as pop

    rec, _ = pop.fly_pid_population_host(paramset="11")
    assert (int(rec["reward"][0]), int(rec["flight_time"][0]), int(rec["coins"][0])) == (5923, 1577, 4)
    assert rec["status"][0] & STATUS_FINISHED
    names = ["00", "01", "02", "10", "11", "12", "13", "20", "21", "22"]
    rec, _ = pop.fly_pid_population_host(paramset=names)  # compare_paramsets.sh: nine GUI runs -> one launch
    assert rec["reward"].tolist() == [6059, 5860, 5198, 5933, 5923, 5990, 5946, 6098, 5910, 5928]
    with pytest.raises(TypeError):
        pop.fly_pid_population_host()
