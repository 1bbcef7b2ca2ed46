"""CPU checks of the oracle's GA restatement (evolve/util.py:5-78, drone/drone.py:167-172 with the
library's Philox stream): structure of children, hyper-parameter statistics, shard consistency."""
import numpy as np

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
