"""Brain container — interface of the reference's evolve/brain.py:9-76.

`weights` is the list of per-layer (n_out, n_in + 1) float64 arrays, bias in the LAST column,
that the reference's GA operators and jsonpickle checkpoints manipulate.  During fly() the
forward pass runs inside the CUDA episode kernel (csrc/ldrones.cu `mlp_mac`), and
`translate_input_to_thrust` of a Drone-shaped brain runs one GPU step; `feed_forward` is a
host-side inspection helper with the reference's semantics (any layer shapes) that
fly()/fly_population() never call.
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
        """(4, 1) thrust for one state.  Brains of the Drone's shape run on the GPU with the episode
        kernel's arithmetic (ld_drone_step_host); other layer shapes use the host helper above."""
        if not self.is_gpu_shape():
            return self.feed_forward(np.vstack((drone_position, drone_velocity, drone_angle, drone_angle_velocity, coin_position)))
        import ctypes as C

        from .. import _lib
        from .._abi import STATE_LEN, make_params

        state = np.zeros((1, STATE_LEN))
        for lo, v in ((0, drone_position), (3, drone_velocity), (6, drone_angle), (9, drone_angle_velocity)):
            state[0, lo:lo + 3] = np.asarray(v, dtype=np.float64).ravel()
        genome = self.genome()[None, :].copy()
        coin = np.asarray(coin_position, dtype=np.float64).reshape(1, 3).copy()
        p = make_params(timesteps=1)
        _lib.check(_lib.load().ld_drone_step_host(C.byref(p), 2, genome.ctypes.data, state.ctypes.data, coin.ctypes.data, None, 1))
        return state[0, 24:28].reshape(4, 1).copy()
