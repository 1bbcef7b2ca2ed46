"""Brain container — interface of the reference's evolve/brain.py:9-76.

`weights` is the list of per-layer (n_out, n_in + 1) float64 arrays, bias in the LAST column,
that the reference's GA operators and jsonpickle checkpoints manipulate.  During fly() the
forward pass runs inside the CUDA episode kernel (csrc/ld_step.cuh `dense`); `feed_forward`
here is a host-side inspection helper with the reference's semantics and is never called by
fly()/fly_population().
"""
import numpy as np
import numpy.random as npr

from .._abi import GENOME_LEN, LAYER_SHAPES


class NeuralNetwork:
    def __init__(self, layer_spec=None, n_inputs=15):
        """layer_spec: {layer_number: n_neurons}; default is the reference's {1: 8, 2: 10, 3: 4}
        (evolve/brain.py:10); Drone uses {1: 10, 2: 10, 3: 4} (drone/drone.py:65)."""
        layer_spec = dict({1: 8, 2: 10, 3: 4} if layer_spec is None else layer_spec)
        self.weights = []
        self.n_layers = len(layer_spec)
        layer_spec[0] = n_inputs
        self.layer_spec = layer_spec
        for layer_num, n_neurons in layer_spec.items():
            if layer_num == 0:
                continue
            layer_inputs = layer_spec[layer_num - 1]
            # N(0, 1/sqrt(n_out)) from the global numpy stream, as evolve/brain.py:24
            self.weights.append(npr.normal(0.0, 1.0 / np.sqrt(n_neurons), size=(n_neurons, layer_inputs + 1)))

    # -- genome <-> weights ------------------------------------------------------------------
    def is_gpu_shape(self) -> bool:
        return tuple(np.shape(w) for w in self.weights) == LAYER_SHAPES

    def genome(self) -> np.ndarray:
        """Flatten like evolve/util.py:58-59 does (layer.flatten(), layer after layer)."""
        if not self.is_gpu_shape():
            raise ValueError(
                f"the episode kernel is built for layer shapes {LAYER_SHAPES}; got "
                f"{tuple(np.shape(w) for w in self.weights)}"
            )
        return np.concatenate([np.asarray(w, dtype=np.float64).flatten() for w in self.weights])

    def set_genome(self, g):
        g = np.asarray(g, dtype=np.float64).reshape(GENOME_LEN)
        out, o = [], 0
        for r, c in LAYER_SHAPES:
            out.append(g[o : o + r * c].reshape(r, c).copy())
            o += r * c
        self.weights = out

    # -- host-side inspection (reference semantics) -------------------------------------------
    def _relu(self, x, thresh=0, alpha=1.0):
        return 0.0 if x < thresh else alpha * x

    def relu(self, x_vec, thresh=0, alpha=1.0):
        return np.array([[self._relu(x, thresh=thresh, alpha=alpha) for x in np.squeeze(x_vec)]]).T

    def sigmoid(self, x, beta=0.4):
        return 1 / (np.exp(-beta * x) + 1)

    def feed_forward(self, input_vec):
        x = input_vec
        for i in range(self.n_layers):
            x = np.dot(self.weights[i], np.vstack((x, [[1]])))
            assert x.shape[1] == 1
            x = self.sigmoid(x) if i == self.n_layers - 1 else self.relu(x)
        return x

    def translate_input_to_thrust(self, drone_position, drone_velocity, drone_angle, drone_angle_velocity, coin_position):
        return self.feed_forward(
            np.vstack((drone_position, drone_velocity, drone_angle, drone_angle_velocity, coin_position))
        )
