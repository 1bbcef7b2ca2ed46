"""Coin target of the coin-collection episode.

Interface-compatible with the reference's objects/coins.py:4-25 (`position`, `x`, `y`, `z`,
`radius`, `value`, `is_in_vicinity`, `distance_drone_to_coin`).  During fly() the test
`||coin - position|| < radius` runs inside the CUDA kernel (csrc/ld_step.cuh,
`step_rewards_and_coin`); this class is the host-side object that scripts and
`datahandling.coin_schedule` use.
"""
import numpy as np

_SHAPE = (3, 1)


def _column(a, what):
    a = np.asarray(a)
    assert a.shape == _SHAPE, f"{what}: expected shape {_SHAPE}, got {a.shape}"
    return a


class Coin:
    __slots__ = ("position", "radius", "value")

    def __init__(self, position, radius, value=100):
        self.position = _column(position, "coin position")
        self.radius = radius
        self.value = value

    # coordinates are views into `position`, so they cannot drift apart
    x = property(lambda self: self.position[0, 0])
    y = property(lambda self: self.position[1, 0])
    z = property(lambda self: self.position[2, 0])

    def distance_drone_to_coin(self, drone_position):
        """Euclidean distance, evaluated like the reference (np.linalg.norm of the difference)."""
        return np.linalg.norm(self.position - drone_position)

    def is_in_vicinity(self, drone_position):
        """Strictly inside the collection radius."""
        return self.distance_drone_to_coin(_column(drone_position, "drone position")) < self.radius
