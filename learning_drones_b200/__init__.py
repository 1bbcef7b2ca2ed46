"""learning_drones_b200 — B200-native population fitness evaluation for learning-drones.

Public surface mirrors the reference's module layout for the hot path:
    learning_drones_b200.flight.fly / fly_population     (reference flight/__init__.py)
    learning_drones_b200.drone.drone.Drone               (reference drone/drone.py)
    learning_drones_b200.evolve.brain.NeuralNetwork      (reference evolve/brain.py)
    learning_drones_b200.evolve.util                     (reference evolve/util.py)
    learning_drones_b200.objects.coins.Coin              (reference objects/coins.py)
Everything numeric on that path runs in libldrones.so (hand-written sm_100a CUDA);
there is no CPU fallback.
"""
__version__ = "0.1.0"
