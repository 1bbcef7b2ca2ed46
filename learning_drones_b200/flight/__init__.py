"""fly() — drop-in for the reference's flight/__init__.py:12-99, plus fly_population().

`fly(drone, ...)` keeps the reference signature, return value `(reward, real_time, idx)` and
in-place mutation of the Drone.  Internally it is a population of one; evolution loops should
call `fly_population(drones, ...)` once per generation instead of N times fly() — same results,
one kernel launch.
"""
import numpy as np

from .._abi import STATUS_NONFINITE, make_params
from ..drone.drone import Drone
from ..population import fly_population_host


def fly_population(drones, timesteps: int = 2500, delta_t: float = 0.01, visualize=False, n_servers=0,
                   punish_time=False, reward_time=True):
    """Fly every Drone of `drones` (one generation, evolve/evolution.py:64-73).

    Returns [(reward, real_time, idx), ...] in population order and mutates the drones exactly
    like the reference's fly() does.
    """
    if visualize or n_servers:
        raise NotImplementedError("the B200 fitness path is headless (visualize=False, n_servers=0)")
    drones = list(drones)
    if not drones:
        return []
    starts = [d.start_state() for d in drones]
    first = starts[0]
    for s in starts[1:]:
        if not all(np.array_equal(a, b) for a, b in zip(first, s)):
            raise ValueError("all drones of one fly_population() call must start from the same state")
    params = make_params(
        timesteps=timesteps, reward_time=reward_time, punish_time=punish_time,
        initial_pos=first[0], initial_vel=first[1], initial_angle=first[2], initial_angle_vel=first[3],
        initial_thrust=first[4],
    )
    genomes = np.stack([d.controller.genome() for d in drones])
    records, state = fly_population_host(genomes, params, want_state=True)
    out = []
    for i, d in enumerate(drones):
        if records[i]["status"] & STATUS_NONFINITE:
            # reference: REWARD_COIN_DISTANCE(nan) -> int(nan) raises (drone/drone.py:14)
            raise ValueError("cannot convert float NaN to integer")
        d.load_result(records[i], state[i])
        if timesteps <= 0:
            # reference: `real_time` is unbound when the loop body never runs (flight/__init__.py:99)
            raise UnboundLocalError("cannot access local variable 'real_time' where it is not associated with a value")
        out.append((d.reward, (d.flight_time - 1) * delta_t, i))
    return out


def fly(drone: Drone, timesteps: int = 2500, delta_t: float = 0.01, visualize=False, n_servers=0,
        punish_time=False, reward_time=True, idx=None):
    reward, real_time, _ = fly_population(
        [drone], timesteps=timesteps, delta_t=delta_t, visualize=visualize, n_servers=n_servers,
        punish_time=punish_time, reward_time=reward_time,
    )[0]
    return reward, real_time, idx
