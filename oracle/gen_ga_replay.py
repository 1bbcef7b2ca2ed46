"""Pin the GA operators to the reference by DRAW REPLAY -> tests/golden/ga_reference_replay.npz.

TEST INFRASTRUCTURE (build container only; needs /root/reference).

The library's GA step (csrc/ldrones.cu `ga_step_kernel`, restated in oracle/ld_oracle.c
`ld_oracle_ga_step`) draws from a counter-based Philox4x32-10 stream, the reference from numpy's
global MT19937, so children cannot be compared under a common seed.  Instead the UNMODIFIED
reference operators (evolve/util.py:5-78 `selection_and_crossover` / `tournament_select` /
`crossover`, drone/drone.py:167-172 `mutate`, :177 `reset`, and the elitism lines of
evolve/evolution.py:77-107) are run here with `np.random.randint / random_sample / normal`
replaced by a replayer that hands out, in the reference's own call order, exactly the draws the
library derives from its Philox counters:

    pair p = child // 2, generation g, key = seed
      randint(0, n) x4          <- floor(u53 * n) of Philox(p, 0|1, SELECT, g)      (util.py:9-12)
      random_sample() x2        <- u53 of Philox(p, 2, SELECT, g)                    (util.py:31, once per tournament)
      per layer l: random_sample() <- u53(x0, x1) of Philox(p, 3 + l, SELECT, g)     (util.py:52)
                   randint(0, genes) <- floor(u53(x2, x3) * genes), only if crossed (util.py:56)
    child c, gene j (C-order flattening, layer after layer):
      random_sample()           <- word (j % 4) of Philox(c, j // 4, MUTATE, g) * 2^-32   (drone.py:171)
      normal(0, 0.3)            <- 0.3 * Box-Muller of Philox(c, j, NOISE, g), only if mutated (drone.py:172)

The replayer ASSERTS the kind and arguments of every call, so the fixture also pins the
reference's draw order.  tests/test_ga_oracle.py then requires `oracle.ga_step` (+ elitism) to
reproduce the reference's children bit for bit, tests/test_gpu_ga.py the CUDA kernel (mutated
genes to 1e-12: CUDA's log/cos differ from glibc's in the last ulps).

    python oracle/gen_ga_replay.py
"""
import contextlib
import io
import math
import os
import sys
import tempfile
from copy import deepcopy

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import gen_golden as G  # noqa: E402

SELECT, MUTATE, NOISE = 1, 2, 3
M32 = 0xFFFFFFFF
LAYER_GENES = (160, 110, 44)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011), python ints."""
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c0, 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0, k1 = (k0 + 0x9E3779B9) & M32, (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def u53(hi, lo):
    return float(((hi << 32) | lo) >> 11) * 2.0 ** -53


class Replayer:
    """Hands the library's Philox draws to the reference in the reference's call order."""

    def __init__(self, seed, generation, n, p_crossover, mutation_rate, sigma):
        self.k0, self.k1, self.gen, self.n = seed & M32, (seed >> 32) & M32, generation, n
        self.p_crossover, self.rate, self.sigma = p_crossover, mutation_rate, sigma
        self.calls = 0
        self.script = self._script()
        self.expect = next(self.script)

    def block(self, a, b, stream):
        return philox4x32_10(a & M32, b & M32, stream, self.gen, self.k0, self.k1)

    def _script(self):
        """Generator of (kind, args, value); receives nothing — the value handed out decides what comes next."""
        n = self.n
        for pair in range(n // 2):
            a, b, t = self.block(pair, 0, SELECT), self.block(pair, 1, SELECT), self.block(pair, 2, SELECT)
            for hi, lo in ((a[0], a[1]), (a[2], a[3]), (b[0], b[1]), (b[2], b[3])):
                yield "randint", (0, n), int(u53(hi, lo) * float(n))
            yield "random_sample", (), u53(t[0], t[1])
            yield "random_sample", (), u53(t[2], t[3])
            for l, genes in enumerate(LAYER_GENES):
                x = self.block(pair, 3 + l, SELECT)
                r = u53(x[0], x[1])
                yield "random_sample", (), r
                if r < self.p_crossover:
                    yield "randint", (0, genes), int(u53(x[2], x[3]) * float(genes))
        for child in range(n):
            m = None
            for j in range(sum(LAYER_GENES)):
                if j % 4 == 0:
                    m = self.block(child, j // 4, MUTATE)
                u = float(m[j % 4]) * 2.0 ** -32
                yield "random_sample", (), u
                if u < self.rate:
                    x = self.block(child, j, NOISE)
                    u1, u2 = (float(x[0]) + 0.5) * 2.0 ** -32, (float(x[1]) + 0.5) * 2.0 ** -32
                    z = math.sqrt(-2.0 * math.log(u1)) * math.cos(6.283185307179586 * u2)
                    yield "normal", (0.0, self.sigma), self.sigma * z
        yield "end", (), None

    def _take(self, kind, args):
        want_kind, want_args, value = self.expect
        assert (kind, tuple(args)) == (want_kind, tuple(want_args)), f"call {self.calls}: reference asked {kind}{args}, script expects {want_kind}{want_args}"
        self.calls += 1
        self.expect = next(self.script)
        return value

    def randint(self, low, high=None, *a, **kw):
        assert not a and not kw
        return self._take("randint", (low, high))

    def random_sample(self, *a, **kw):
        assert not a and not kw
        return self._take("random_sample", ())

    def normal(self, loc=0.0, scale=1.0, *a, **kw):
        assert not a and not kw
        return self._take("normal", (loc, scale))

    def finished(self):
        return self.expect[0] == "end"


@contextlib.contextmanager
def patched_numpy_random(rep):
    saved = np.random.randint, np.random.random_sample, np.random.normal
    np.random.randint, np.random.random_sample, np.random.normal = rep.randint, rep.random_sample, rep.normal
    try:
        yield
    finally:
        np.random.randint, np.random.random_sample, np.random.normal = saved


def reference_generation_step(Drone, rutil, genomes, rewards, rep, p_tournament, p_crossover, mutation_rate, global_best=None):
    """evolve/evolution.py:77-107 on reference objects: argmax, global best (deepcopy), selection_and_crossover,
    reset + mutate of every child, elitism.  `global_best` = (score, genome) carried in from earlier generations."""
    with contextlib.redirect_stdout(io.StringIO()):
        drones = [Drone(visualize=False, n_servers=0) for _ in genomes]
    for d, g, r in zip(drones, genomes, rewards):
        d.controller.weights = G.split_genome(g)
        d.reward = int(r)
    scores = [int(r) for r in rewards]
    best_index = int(np.argmax(scores))                                   # :77
    best_score = scores[best_index]
    global_best_score, global_best_drone = (-np.inf, None)
    if global_best is not None:
        with contextlib.redirect_stdout(io.StringIO()):
            global_best_drone = Drone(visualize=False, n_servers=0)
        global_best_drone.controller.weights = G.split_genome(global_best[1])
        global_best_score = global_best[0]
    if best_score > global_best_score:                                    # :80-83
        global_best_score, global_best_drone = best_score, deepcopy(drones[best_index])
    with patched_numpy_random(rep):
        new_population = rutil.selection_and_crossover(drones, p_tournament, p_crossover)   # :89
        for d in new_population:                                          # :91-93
            d.reset()
            d.mutate(mutation_rate)
    assert rep.finished(), "the reference stopped drawing before the script ended"
    before_elitism = np.array([G.flatten_weights(d.controller.weights) for d in new_population])
    drones = new_population                                               # :105-107
    drones[0] = global_best_drone
    drones[0].reset()
    children = np.array([G.flatten_weights(d.controller.weights) for d in drones])
    return children, before_elitism, best_index, global_best_score, G.flatten_weights(global_best_drone.controller.weights)


CASES = (
    # name, n, seed, generation, p_tournament, p_crossover, mutation_rate
    ("default20", 20, 0xC0FFEE12345, 0, 0.6, 0.05, 0.05),      # evolve/evolution.py:16-20 hyper-parameters
    ("crossed20", 20, 77, 3, 0.6, 0.5, 0.05),                  # many crossovers
    ("default256", 256, 2026, 7, 0.6, 0.05, 0.05),
    ("heavy64", 64, (1 << 40) + 5, 1, 0.3, 1.0, 0.25),         # every layer crossed, worse parent favoured
)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.chdir(tempfile.mkdtemp(prefix="ld_ga_replay_"))
    Drone, _ = G._import_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        import evolve.util as rutil
    out = {"cases": np.array([c[0] for c in CASES])}
    for name, n, seed, gen, pt, pc, rate in CASES:
        rng = np.random.default_rng(n + gen)
        genomes = G.random_brains(n, rng)
        rewards = rng.integers(-6000, 4000, size=n)
        rewards[rng.integers(0, n, size=n // 4)] = rewards[0]            # ties: `score1 > score2` is strict (util.py:35-45)
        carried = None
        if name == "default256":                                           # an earlier generation's best outlives this one
            carried = (int(rewards.max()) + 10, G.random_brains(1, rng)[0])
        rep = Replayer(seed, gen, n, pc, rate, 0.3)
        children, before, best_index, gbest, gbest_genome = reference_generation_step(Drone, rutil, genomes, rewards, rep, pt, pc, rate, carried)
        out.update({f"{name}_parents": genomes, f"{name}_rewards": rewards.astype(np.int64), f"{name}_children": children,
                    f"{name}_children_before_elitism": before, f"{name}_best_index": np.int64(best_index),
                    f"{name}_global_best_score": np.int64(gbest), f"{name}_global_best_genome": gbest_genome,
                    f"{name}_hyper": np.array([float(seed), gen, pt, pc, rate, 0.3]), f"{name}_seed": np.uint64(seed),
                    f"{name}_carried_score": np.int64(carried[0] if carried else -(2 ** 62)),
                    f"{name}_carried_genome": carried[1] if carried else np.zeros(G.GENOME_LEN), f"{name}_draws": np.int64(rep.calls)})
        changed = (before != np.zeros_like(before)).sum()
        print(f"{name}: n={n} draws={rep.calls} best_index={best_index} global_best={gbest} nonzero genes={changed}")
    np.savez_compressed(os.path.join(out_dir, "ga_reference_replay.npz"), **out)
    print("written", os.path.join(out_dir, "ga_reference_replay.npz"))


if __name__ == "__main__":
    main()
