"""CPU checks of the oracle's GA restatement (evolve/util.py:5-78, drone/drone.py:167-172 with the
library's Philox stream): structure of children, hyper-parameter statistics, shard consistency."""
import numpy as np
import pytest

from learning_drones_b200._abi import GENOME_LEN, make_ga_params

LAYERS = ((0, 160), (160, 270), (270, 314))


def test_oracle_ga_structure_and_statistics(oracle):
    n = 6000
    parents = np.tile(np.arange(n, dtype=np.float64)[:, None], (1, GENOME_LEN))
    rewards = np.arange(n, dtype=np.int64)
    ga = make_ga_params(seed=12, generation=3, mutation_rate=0.0)
    kids, picks = oracle.ga_step(ga, parents, rewards)
    assert np.array_equal(kids[:, 0].astype(np.int64), picks[:, 0])
    assert abs(picks[:, 0].mean() / n - (0.6 * 2 / 3 + 0.4 / 3)) < 0.02  # tournament pressure, util.py:30-45
    for l, (a, b) in enumerate(LAYERS):
        cut = picks[:, 1 + l]
        assert abs((cut >= 0).mean() - 0.05) < 0.012 and cut.max() < b - a
        assert (cut[0::2] == cut[1::2]).all()  # one draw per pair and layer, util.py:48-57
        crossed = np.nonzero(cut >= 0)[0]
        for c in crossed[:50]:
            assert (kids[c, a : a + cut[c]] == picks[c, 0]).all() and (kids[c, a + cut[c] : b] == picks[c ^ 1, 0]).all()
    # mutation: Bernoulli(rate) mask, N(0, sigma) noise
    ga = make_ga_params(seed=12, generation=4, p_crossover=0.0, mutation_rate=0.05, mutation_sigma=0.3)
    base = np.zeros((n, GENOME_LEN))
    kids, picks = oracle.ga_step(ga, base, np.zeros(n, dtype=np.int64))
    m = kids != 0
    assert m.sum(axis=1).tolist() == picks[:, 4].tolist()
    assert abs(m.mean() - 0.05) < 0.002 and abs(kids[m].std() - 0.3) < 0.01
    # shards and generations
    a, _ = oracle.ga_step(ga, base, np.zeros(n, dtype=np.int64), 100, 300)
    assert np.array_equal(a, kids[100:300])
    b, _ = oracle.ga_step(make_ga_params(seed=12, generation=5, p_crossover=0.0), base, np.zeros(n, dtype=np.int64))
    assert not np.array_equal(b, kids)


REPLAY_CASES = ("default20", "crossed20", "default256", "heavy64")


def replay_case(golden, name):
    """One case of tests/golden/ga_reference_replay.npz: the REFERENCE's selection_and_crossover + reset + mutate +
    elitism (evolve/util.py:5-78, drone/drone.py:167-172, evolve/evolution.py:77-107) driven with the library's
    Philox draws in the reference's own call order (oracle/gen_ga_replay.py)."""
    g = golden("ga_reference_replay")
    _, gen, pt, pc, rate, sigma = g[f"{name}_hyper"].tolist()
    ga = make_ga_params(seed=int(g[f"{name}_seed"]), generation=int(gen), p_tournament=pt, p_crossover=pc, mutation_rate=rate,
                        mutation_sigma=sigma)
    return g, ga


@pytest.mark.parametrize("name", REPLAY_CASES)
def test_oracle_ga_reproduces_reference_children(oracle, golden, name):
    g, ga = replay_case(golden, name)
    parents, rewards = g[f"{name}_parents"], g[f"{name}_rewards"]
    kids, picks = oracle.ga_step(ga, parents, rewards)
    # children of the reference's operators, bit for bit (same glibc log/cos/sqrt in the Box-Muller draw)
    assert np.array_equal(kids, g[f"{name}_children_before_elitism"])
    assert int(g[f"{name}_draws"]) > 0 and picks[:, 4].sum() > 0
    if name == "crossed20":
        assert (picks[:, 1:4] >= 0).sum() >= 6  # the case exists to exercise crossovers
    # evolution.py:77-83,105-107: best so far (first arg-max, strictly better than the carried one) goes to slot 0
    best = int(np.argmax(rewards))
    assert best == int(g[f"{name}_best_index"])
    carried = int(g[f"{name}_carried_score"])
    elite = parents[best] if int(rewards[best]) > carried else g[f"{name}_carried_genome"]
    assert np.array_equal(elite, g[f"{name}_global_best_genome"])
    kids[0] = elite
    assert np.array_equal(kids, g[f"{name}_children"])
