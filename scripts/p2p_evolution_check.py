"""torchrun check (>= 2 GPUs): the p2p-sharded evolution loop gives bit-identical populations and
histories to the replicated one (dev/CI tool; run under gpurun --gpus 2)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from learning_drones_b200.evolve.gpu import GpuEvolution

torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
rng = np.random.default_rng(3)
genomes = np.concatenate([rng.normal(0.0, 1.0 / np.sqrt(r), size=(4096, r * c)) for r, c in ((10, 16), (10, 11), (4, 11))], axis=1)
a = GpuEvolution(genomes, seed=5, timesteps=500, p2p=True)
b = GpuEvolution(genomes, seed=5, timesteps=500, p2p=False)
ha, hb = a.run(6), b.run(6)
pa, pb = a.gather_population(), b.gather_population()
ok = ha == hb and np.array_equal(pa, pb) and np.array_equal(a.best_genome(), b.best_genome()) and a.p2p and not b.p2p
# the same with every rank passing only its own shard (big populations: bench.py's config 5), both GA modes
lo, hi = a.lo, a.hi
c = GpuEvolution(genomes[lo:hi], seed=5, timesteps=500, p2p=True, local_shard=True, n_global=len(genomes))
d = GpuEvolution(genomes[lo:hi], seed=5, timesteps=500, p2p=False, local_shard=True, n_global=len(genomes))
hc, hd = c.run(6), d.run(6)
ok = ok and hc == ha and hd == ha and np.array_equal(c.gather_population(), pa) and np.array_equal(d.gather_population(), pa)
ok = ok and (a.ga_mode, b.ga_mode, c.ga_mode, d.ga_mode) == ("p2p", "replicated", "p2p", "replicated")
if dist.get_rank() == 0:
    print("p2p == replicated:", ok, "best", ha[-1]["global_best_score"], [h["best_score"] for h in ha])
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
