"""Coin target — same interface as the reference's objects/coins.py:4-25."""
import numpy as np


class Coin:
    def __init__(self, position, radius, value=100):
        """position: (3, 1) array"""
        assert position.shape == (3, 1), position.shape
        self.position = position
        self.x, self.y, self.z = position[0, 0], position[1, 0], position[2, 0]
        self.radius = radius
        self.value = value

    def distance_drone_to_coin(self, drone_position):
        return np.linalg.norm(self.position - drone_position)

    def is_in_vicinity(self, drone_position):
        assert drone_position.shape == (3, 1), drone_position.shape
        return self.distance_drone_to_coin(drone_position) < self.radius
