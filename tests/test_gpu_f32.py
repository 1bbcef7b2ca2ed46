"""fp32 fast mode: statistical agreement with the fp64 parity kernel (no parity claim)."""
import numpy as np
import pytest

from learning_drones_b200._abi import make_params

pytestmark = pytest.mark.gpu


def test_f32_survivors_agree_with_f64():
    """The (S) family is not chaotic (pure vertical motion): float32 must land within float32 accuracy."""
    from oracle import gen_golden as G

    from learning_drones_b200 import population as pop

    g = G.survivors(8192, np.random.default_rng(2026))
    p = make_params(2000)
    r64, s64 = pop.fly_population_host(g, p, want_state=True)
    r32, s32 = pop.fly_population_host(g, p, want_state=True, precision="f32")
    assert (r32["flight_time"] == 2000).all() and (r32["status"] == 0).all()
    assert np.abs(s32[:, 2] - s64[:, 2]).max() < 2e-2 * (1 + np.abs(s64[:, 2]).max())  # z after 2,000 steps
    assert np.median(np.abs(s32[:, 2] - s64[:, 2])) < 1e-3
    # reward = 4000 - 100*floor(dist): equal unless the distance sits next to an integer
    assert (r32["reward"] == r64["reward"]).mean() > 0.97
    assert np.abs(r32["reward"] - r64["reward"]).max() <= 100


def test_f32_random_population_statistics():
    from oracle import gen_golden as G

    from learning_drones_b200 import population as pop

    rng = np.random.default_rng(2026)
    G.survivors(65536, rng)
    g = G.random_brains(65536, rng)[:20000]
    p = make_params(2500)
    r64, _ = pop.fly_population_host(g, p)
    r32, _ = pop.fly_population_host(g, p, precision="f32")
    # same physics: distributions agree, individual (chaotic) drones need not
    assert abs(r32["flight_time"].mean() / r64["flight_time"].mean() - 1) < 0.01
    assert abs(np.median(r32["reward"]) - np.median(r64["reward"])) <= 30
    assert abs(int(r32["coins"].sum()) - int(r64["coins"].sum())) <= 0.1 * max(int(r64["coins"].sum()), 20)
    assert ((r32["status"] & 1) == (r64["status"] & 1)).mean() > 0.999
    same = (r32["reward"] == r64["reward"])
    assert 0.3 < same.mean() <= 1.0  # ~57 % of random drones keep their integer reward in float32
    dt = np.abs(r32["flight_time"] - r64["flight_time"])
    assert (dt <= 2).mean() > 0.7 and (dt <= 25).mean() > 0.95
    # deterministic and permutation invariant like the fp64 mode
    r32b, _ = pop.fly_population_host(g, p, precision="f32")
    assert r32b.tobytes() == r32.tobytes()
    # ragged / edge sizes and the device entry point
    import torch

    for n in (0, 1, 255, 257):
        a, _ = pop.fly_population_host(g[:n], p, precision="f32")
        assert a.tobytes() == r32[:n].tobytes()
    rd, _ = pop.fly_population_device(torch.from_numpy(g[:5000]).cuda(), p, precision="f32")
    assert pop.records_to_numpy(rd).tobytes() == r32[:5000].tobytes()
    with pytest.raises(ValueError):
        pop.fly_population_host(g[:4], p, precision="bf16")
