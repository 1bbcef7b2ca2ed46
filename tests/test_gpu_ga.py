"""GA step on the GPU: kernel vs the oracle's restatement with the same Philox
stream (indices/cuts/masks exact, mutated values to 1e-12), semantics of evolve/util.py, and the
device-resident evolution loop."""
import ctypes as C

import numpy as np
import pytest

from learning_drones_b200._abi import GENOME_LEN, RECORD_DTYPE, make_ga_params

pytestmark = pytest.mark.gpu
LAYERS = ((0, 160), (160, 270), (270, 314))


def gpu_ga(ga, parents, rewards, lo=0, hi=None):
    import torch

    from learning_drones_b200 import _lib

    L = _lib.load()
    n = parents.shape[0]
    hi = n if hi is None else hi
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    rec["reward"] = rewards
    pt = torch.from_numpy(parents).cuda()
    rt = torch.from_numpy(rec.view(np.int64).reshape(-1, 4).copy()).cuda()
    out = torch.zeros((hi - lo, GENOME_LEN), dtype=torch.float64, device="cuda")
    _lib.check(L.ld_ga_step_device(C.byref(ga), pt.data_ptr(), rt.data_ptr(), n, out.data_ptr(), lo, hi, None))
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("n,pc,rate", [(20, 0.05, 0.05), (4096, 0.5, 0.05), (1000, 1.0, 0.0), (64, 0.0, 1.0)])
def test_ga_kernel_matches_oracle(oracle, n, pc, rate):
    rng = np.random.default_rng(n)
    parents = rng.standard_normal((n, GENOME_LEN))
    rewards = rng.integers(-5000, 4000, size=n)
    ga = make_ga_params(seed=0xC0FFEE12345, generation=7, p_crossover=pc, mutation_rate=rate)
    want, picks = oracle.ga_step(ga, parents, rewards)
    got = gpu_ga(ga, parents, rewards)
    assert np.abs(got - want).max() < 1e-12  # Box-Muller goes through log/cos: libm vs CUDA ulps
    mutated = np.ones_like(want, dtype=bool)
    for c in range(n):
        pair = c ^ 1
        for l, (a, b) in enumerate(LAYERS):
            cut = picks[c, 1 + l]
            src_head, src_tail = picks[c, 0], picks[pair, 0] if cut >= 0 else picks[c, 0]
            cutpos = a + (cut if cut >= 0 else (b - a))
            base = np.concatenate([parents[src_head, a:cutpos], parents[src_tail, cutpos:b]])
            mutated[c, a:b] = want[c, a:b] != base
            assert np.array_equal(got[c, a:b][~mutated[c, a:b]], base[~mutated[c, a:b]])  # bit-exact copies
    assert mutated.sum(axis=1).tolist() == picks[:, 4].tolist()
    # shard form: rows lo..hi of the same generation are identical to the full call
    lo, hi = (n // 4) * 2, (n // 4) * 2 + min(n // 2, 10) // 2 * 2
    assert np.array_equal(gpu_ga(ga, parents, rewards, lo, hi), got[lo:hi])


def test_ga_semantics():
    """Statistics of evolve/util.py / Drone.mutate: tournament pressure, per-layer crossover rate, mutation N(0, 0.3)."""
    rng = np.random.default_rng(5)
    n = 20000
    parents = np.tile(np.arange(n, dtype=np.float64)[:, None], (1, GENOME_LEN))  # gene value = parent id
    rewards = np.arange(n)  # reward rank = id
    ga = make_ga_params(seed=99, generation=0, p_tournament=0.6, p_crossover=0.05, mutation_rate=0.0)
    kids = gpu_ga(ga, parents, rewards)
    head = kids[:, 0].astype(np.int64)  # parent each child copied its first gene from
    # winner of a 2-way tournament with p=0.6 for the better: E[rank/n] = 0.6*2/3 + 0.4*1/3 = 0.5333
    assert abs(head.mean() / n - (0.6 * 2 / 3 + 0.4 / 3)) < 0.01
    crossed = [(kids[:, a] != kids[:, b - 1]) for a, b in LAYERS]
    for c in crossed:  # a crossed layer shows two different parent ids unless cut == 0 or both parents equal
        assert 0.03 < c.mean() < 0.06
    # each child layer is a prefix of one parent and a suffix of the pair partner
    for a, b in LAYERS:
        seg = kids[:, a:b]
        changes = (np.diff(seg, axis=1) != 0).sum(axis=1)
        assert changes.max() <= 1
    ga = make_ga_params(seed=3, generation=1, p_crossover=0.0, mutation_rate=0.05, mutation_sigma=0.3)
    base = rng.standard_normal((n, GENOME_LEN))
    kids = gpu_ga(ga, base, np.zeros(n, dtype=np.int64))  # equal rewards: ties -> second index (util.py:35-45)
    # find each child's parent through an unmutated gene pattern is overkill: use the noise statistics
    ga0 = make_ga_params(seed=3, generation=1, p_crossover=0.0, mutation_rate=0.0)
    noise = kids - gpu_ga(ga0, base, np.zeros(n, dtype=np.int64))
    m = noise != 0
    assert abs(m.mean() - 0.05) < 0.002
    assert abs(noise[m].std() - 0.3) < 0.005 and abs(noise[m].mean()) < 0.005
    with pytest.raises(Exception):
        gpu_ga(ga, base[:7], np.zeros(7, dtype=np.int64))  # odd population (util.py:17 would index out of range)


def test_gpu_evolution_loop():
    from oracle import gen_golden as G

    from learning_drones_b200.evolve.gpu import GpuEvolution

    rng = np.random.default_rng(8)
    genomes = G.random_brains(2000, rng)
    ev = GpuEvolution(genomes, seed=1, timesteps=600)
    hist = ev.run(5)
    g = [h["global_best_score"] for h in hist]
    assert all(b >= a for a, b in zip(g, g[1:]))  # elitism keeps the all-time best
    assert all(h["best_score"] <= h["global_best_score"] for h in hist)
    ev2 = GpuEvolution(genomes, seed=1, timesteps=600)
    assert ev2.run(5) == hist  # counter-based RNG: reproducible
    assert not np.array_equal(ev.population.cpu().numpy(), genomes)
    assert np.array_equal(ev.population[0].cpu().numpy(), ev.best_genome())
    assert ev.global_best_score == hist[-1]["global_best_score"]
    # generation by generation (one host synchronisation each) == enqueued in one block
    ev3 = GpuEvolution(genomes, seed=1, timesteps=600)
    assert [ev3.step() for _ in range(5)] == hist
    assert np.array_equal(ev3.population.cpu().numpy(), ev.population.cpu().numpy())
    # device-side best tracking == host-side: the best genome is the row that produced the best reward
    from learning_drones_b200.population import fly_population_device, records_to_numpy

    ev4 = GpuEvolution(genomes, seed=1, timesteps=600)
    best, best_row = None, None
    for _ in range(3):
        pop_before = ev4.population.clone()
        h = ev4.step()
        r = records_to_numpy(fly_population_device(pop_before, ev4.params)[0])["reward"]
        assert h["best_score"] == int(r.max()) and h["best_index"] == int(np.argmax(r))
        if best is None or h["best_score"] > best:
            best, best_row = h["best_score"], pop_before[h["best_index"]].cpu().numpy()
        assert h["global_best_score"] == best and np.array_equal(ev4.best_genome(), best_row)


# ---- pinned to the reference by draw replay (oracle/gen_ga_replay.py) ---------------------------------
from test_ga_oracle import REPLAY_CASES, replay_case  # noqa: E402


@pytest.mark.parametrize("name", REPLAY_CASES)
def test_ga_kernel_reproduces_reference_children(golden, name):
    """The CUDA GA step + device-side best tracking + elitism against children produced by the REFERENCE's own
    selection_and_crossover / reset / mutate / elitism code fed with the library's draws."""
    import torch

    from learning_drones_b200 import _lib
    from learning_drones_b200.population import reduce_records_device

    L = _lib.load()
    g, ga = replay_case(golden, name)
    parents, rewards = g[f"{name}_parents"], g[f"{name}_rewards"]
    n = parents.shape[0]
    want = g[f"{name}_children_before_elitism"]
    got = gpu_ga(ga, parents, rewards)
    # a gene the reference did not mutate is a bit-exact copy of a parent gene; mutated genes carry CUDA's log/cos
    is_copy = np.zeros_like(want, dtype=bool)
    for c in range(n):
        is_copy[c] = (want[c][None, :] == parents).any(axis=0)
    assert np.array_equal(got[is_copy], want[is_copy])
    assert np.abs(got - want).max() < 1e-12 and (~is_copy).sum() > 0
    # evolution.py:77-107 on the device: arg-max, best-so-far (possibly carried in from an earlier generation), slot 0
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    rec["reward"] = rewards
    rt = torch.from_numpy(rec.view(np.int64).reshape(-1, 4).copy()).cuda()
    pt = torch.from_numpy(parents).cuda()
    carried = int(g[f"{name}_carried_score"])
    best2 = torch.tensor([carried, 1] if carried > -(2 ** 61) else [0, 0], dtype=torch.int64, device="cuda")
    best_genome = torch.from_numpy(g[f"{name}_carried_genome"].copy()).cuda()
    hist = torch.zeros(7, dtype=torch.int64, device="cuda")
    stats = reduce_records_device(rt)
    _lib.check(L.ld_track_best_device(stats.data_ptr(), pt.data_ptr(), None, int(ga.generation), best2.data_ptr(), best_genome.data_ptr(),
                                      hist.data_ptr(), None))
    kids = torch.from_numpy(got).cuda()
    kids[0].copy_(best_genome)
    torch.cuda.synchronize()
    assert hist.tolist()[2] == int(g[f"{name}_best_index"]) and best2.tolist() == [int(g[f"{name}_global_best_score"]), 1]
    assert np.array_equal(best_genome.cpu().numpy(), g[f"{name}_global_best_genome"])
    final = kids.cpu().numpy()
    assert np.array_equal(final[0], g[f"{name}_children"][0])
    assert np.abs(final - g[f"{name}_children"]).max() < 1e-12


# ---- the peer-memory instantiation (ga_step_kernel<true>, track_best over a peer table) on ONE GPU ----
@pytest.mark.parametrize("world,separate", [(2, False), (4, True), (8, False), (8, True)])
def test_ga_p2p_kernel_with_local_peer_table(world, separate):
    """ld_ga_step_p2p_device / ld_track_best_device(peers) address parents through a table of per-rank shard
    pointers (NVLink peer mappings on a multi-GPU box).  Here the "peers" are shards in this GPU's own memory —
    slices of one buffer, or separately allocated tensors — so the P2P code path runs wherever the tests run;
    every rank's children must be bit-identical to the single-buffer kernel's."""
    import torch

    from learning_drones_b200 import _lib
    from learning_drones_b200._abi import make_peer_table
    from learning_drones_b200.population import reduce_records_device

    L = _lib.load()
    rows = 250
    n = world * rows
    rng = np.random.default_rng(world)
    parents = rng.standard_normal((n, GENOME_LEN))
    rewards = rng.integers(-5000, 4000, size=n)
    ga = make_ga_params(seed=0xABCDEF987, generation=11, p_crossover=0.3, mutation_rate=0.05)
    want = gpu_ga(ga, parents, rewards)
    rec = np.zeros(n, dtype=RECORD_DTYPE)
    rec["reward"] = rewards
    rt = torch.from_numpy(rec.view(np.int64).reshape(-1, 4).copy()).cuda()
    pt = torch.from_numpy(parents).cuda()
    if separate:
        shards = [pt[r * rows:(r + 1) * rows].clone() for r in reversed(range(world))][::-1]  # allocation order != rank order
    else:
        shards = [pt[r * rows:(r + 1) * rows] for r in range(world)]
    table = make_peer_table([s.data_ptr() for s in shards], rows)
    for r in range(world):
        out = torch.zeros((rows, GENOME_LEN), dtype=torch.float64, device="cuda")
        _lib.check(L.ld_ga_step_p2p_device(C.byref(ga), C.byref(table), rt.data_ptr(), n, out.data_ptr(), r * rows, (r + 1) * rows, None))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), want[r * rows:(r + 1) * rows]), f"rank {r}"
    # best tracking through the peer table == through the whole population
    stats = reduce_records_device(rt)
    outs = []
    for peers in (None, table):
        best2 = torch.zeros(2, dtype=torch.int64, device="cuda")
        best_genome = torch.zeros(GENOME_LEN, dtype=torch.float64, device="cuda")
        hist = torch.zeros(7, dtype=torch.int64, device="cuda")
        _lib.check(L.ld_track_best_device(stats.data_ptr(), pt.data_ptr() if peers is None else None,
                                          None if peers is None else C.byref(peers), 3, best2.data_ptr(), best_genome.data_ptr(),
                                          hist.data_ptr(), None))
        torch.cuda.synchronize()
        outs.append((best2.tolist(), best_genome.cpu().numpy(), hist.tolist()))
    assert outs[0][0] == outs[1][0] == [int(rewards.max()), 1] and outs[0][2] == outs[1][2]
    assert np.array_equal(outs[0][1], outs[1][1]) and np.array_equal(outs[0][1], parents[int(np.argmax(rewards))])
    # argument validation of the peer form
    bad = make_peer_table([s.data_ptr() for s in shards], rows + 2)
    out = torch.zeros((rows, GENOME_LEN), dtype=torch.float64, device="cuda")
    assert L.ld_ga_step_p2p_device(C.byref(ga), C.byref(bad), rt.data_ptr(), n, out.data_ptr(), 0, rows, None) != 0


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_evolution_loop_on_one_gpu(world):
    """GpuEvolution's sharded loop (fly shard -> gather records -> track best over peers -> P2P GA step -> elitism
    on rank 0) run for `world` emulated ranks on one GPU must give the single-rank loop's populations."""
    from oracle import gen_golden as G

    from learning_drones_b200.evolve.gpu import GpuEvolution, LocalShardedEvolution

    rng = np.random.default_rng(4)
    genomes = G.random_brains(world * 256, rng)
    a = GpuEvolution(genomes, seed=9, timesteps=400)
    b = LocalShardedEvolution(genomes, world=world, seed=9, timesteps=400)
    assert a.ga_mode == "single" and b.ga_mode == "p2p-local"
    ha, hb = a.run(4), b.run(4)
    assert ha == hb
    assert np.array_equal(a.gather_population(), b.gather_population())
    assert np.array_equal(a.best_genome(), b.best_genome())


def test_evolution_schedule_adapts_and_never_changes_results():
    """GpuEvolution flies without episode segments while no drone of the last generation it read back survived, with the
    doubling schedule otherwise (evolve/gpu.py, _adapt_schedule) — scheduling only: the generations are bit-identical to
    a run with a fixed schedule."""
    from learning_drones_b200.evolve.gpu import GpuEvolution

    rng = np.random.default_rng(5)
    n, H = 4096, 600
    young = np.concatenate([rng.normal(0.0, 1.0 / np.sqrt(r), size=(n, r * c)) for r, c in ((10, 16), (10, 11), (4, 11))], axis=1)
    adaptive = GpuEvolution(young, seed=3, timesteps=H)
    fixed = GpuEvolution(young, seed=3, timesteps=H, slice_steps=-150)
    assert adaptive.params.slice_steps == -150  # doubling from H/4 until something is known about the population
    a = adaptive.run(2) + adaptive.run(3)
    b = fixed.run(2) + fixed.run(3)
    assert a == b
    assert np.array_equal(adaptive.population.cpu().numpy(), fixed.population.cpu().numpy())
    assert fixed.params.slice_steps == -150
    assert adaptive.params.slice_steps == (-1 if a[-1]["crashed"] >= n else -150)
    # a population with survivors switches the segments back on
    hover = np.zeros((n, GENOME_LEN))
    hover[:, 270 + 10::11] = np.log(0.1962 / 0.8038) / 0.4  # output-layer bias column: hover thrust
    adaptive2 = GpuEvolution(hover, seed=3, timesteps=H)
    adaptive2.params.slice_steps = -1
    h = adaptive2.run(1)
    assert h[-1]["crashed"] < n and adaptive2.params.slice_steps == -150
