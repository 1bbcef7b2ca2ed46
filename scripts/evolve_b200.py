#!/usr/bin/env python
"""GPU counterpart of the reference's `python evolve/evolution.py` (evolve/evolution.py:33-118).

    python scripts/evolve_b200.py --drones 20 --generations 20
    torchrun --nproc-per-node 8 scripts/evolve_b200.py --drones 1048576 --generations 100 --timesteps 2000
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--drones", type=int, default=20)        # n_drones, evolution.py:17
    ap.add_argument("--generations", type=int, default=20)   # n_generations, :16
    ap.add_argument("--timesteps", type=int, default=2500)   # fly() default
    ap.add_argument("--mutation-rate", type=float, default=0.05)
    ap.add_argument("--p-tournament", type=float, default=0.6)
    ap.add_argument("--p-crossover", type=float, default=0.05)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--report-every", type=int, default=10, help="generations enqueued between two host synchronisations")
    ap.add_argument("--out", default=None, help="write the winning genome (314 float64) to this .npy")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from learning_drones_b200.evolve.gpu import GpuEvolution

    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    rank = dist.get_rank() if world > 1 else 0
    # initial brains ~ N(0, 1/sqrt(n_out)) per layer like NeuralNetwork.__init__ (evolve/brain.py:24), same on every rank
    rng = np.random.default_rng(args.seed)
    genomes = np.concatenate([rng.normal(0.0, 1.0 / np.sqrt(r), size=(args.drones, r * c)) for r, c in ((10, 16), (10, 11), (4, 11))], axis=1)
    ev = GpuEvolution(genomes, seed=args.seed, timesteps=args.timesteps, mutation_rate=args.mutation_rate,
                      p_tournament=args.p_tournament, p_crossover=args.p_crossover)
    t0 = time.time()
    done = 0
    while done < args.generations:  # generations are enqueued in blocks; the host only waits to print a block
        block = min(args.report_every, args.generations - done)
        for h in ev.run(block):
            if rank == 0:
                print(f"Generation {h['generation']}: Best Score: {h['best_score']}, Global Best: {h['global_best_score']}  "
                      f"(coins {h['coins']}, crashed {h['crashed']}/{args.drones}, {h['active_steps']} drone-steps)")
        done += block
    torch.cuda.synchronize()
    if rank == 0:
        dt = time.time() - t0
        print(f"{args.generations} generations of {args.drones} drones in {dt:.3f} s = {args.generations / dt:.2f} generations/s")
        if args.out:
            np.save(args.out, ev.best_genome())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
