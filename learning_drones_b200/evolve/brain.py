"""Brain container — interface of the reference's evolve/brain.py:9-76.

`weights` is the list of per-layer (n_out, n_in + 1) float64 arrays, bias in the LAST column,
that the reference's GA operators and jsonpickle checkpoints manipulate.  During fly() the
forward pass runs inside the CUDA episode kernel (csrc/ldrones.cu `mlp_mac`); single calls of
`feed_forward` / `translate_input_to_thrust` / `relu` / `sigmoid` run on the GPU too, for any
layer_spec (`ld_brain_forward_host`, `ld_activation_host`) — nothing here computes on the CPU.
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

    # -- forward pass and activations: always on the GPU (libldrones.so), for ANY layer_spec ------------------
    def _layer_sizes(self):
        return [int(np.shape(self.weights[0])[1]) - 1] + [int(np.shape(w)[0]) for w in self.weights]

    def _relu(self, x, thresh=0, alpha=1.0):
        """brain.py:26-30 for one scalar."""
        return float(self.relu(np.array([[float(x)]]), thresh=thresh, alpha=alpha)[0, 0])

    def _activation(self, kind, x, a, b):
        import ctypes as C

        from .. import _lib

        v = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        out = np.empty_like(v)
        _lib.check(_lib.load().ld_activation_host(kind, v.ctypes.data, v.size, out.ctypes.data, C.c_double(a), C.c_double(b)))
        return out

    def relu(self, x_vec, thresh=0, alpha=1.0):
        """brain.py:32-35: (n, 1) column of `0.0 if x < thresh else alpha * x`."""
        return self._activation(0, np.squeeze(x_vec), float(thresh), float(alpha)).reshape(-1, 1)

    def sigmoid(self, x, beta=0.4):
        """brain.py:37-39: 1 / (exp(-beta x) + 1), same shape as x."""
        return self._activation(1, x, float(beta), 0.0).reshape(np.shape(x))

    def feed_forward(self, input_vec):
        """brain.py:41-54 for whatever layer_spec this network has: bias appended LAST, hidden ReLU, output
        sigmoid.  Runs in libldrones.so (ld_brain_forward_host) with the episode kernel's arithmetic."""
        import ctypes as C

        from .. import _lib

        sizes = self._layer_sizes()
        x = np.ascontiguousarray(np.asarray(input_vec, dtype=np.float64).reshape(1, -1))
        if x.shape[1] != sizes[0]:
            raise ValueError(f"input of {x.shape[1]} values for a network with {sizes[0]} inputs")
        for w, n_in, n_out in zip(self.weights, sizes[:-1], sizes[1:]):
            if np.shape(w) != (n_out, n_in + 1):
                raise ValueError(f"layer of shape {np.shape(w)} does not follow a layer with {n_in} outputs")
        flat = np.ascontiguousarray(np.concatenate([np.asarray(w, dtype=np.float64).reshape(-1) for w in self.weights]))
        out = np.empty((1, sizes[-1]))
        spec = (C.c_int32 * len(sizes))(*sizes)
        _lib.check(_lib.load().ld_brain_forward_host(spec, len(sizes) - 1, flat.ctypes.data, 0, x.ctypes.data, 1, out.ctypes.data, 0.4))
        return out.reshape(-1, 1)

    def translate_input_to_thrust(self, drone_position, drone_velocity, drone_angle, drone_angle_velocity, coin_position):
        """brain.py:56-76: (n_out, 1) thrust for one state — stacks the 15 inputs and feeds them forward (GPU)."""
        return self.feed_forward(np.vstack([np.asarray(v, dtype=np.float64).reshape(3, 1) for v in
                                            (drone_position, drone_velocity, drone_angle, drone_angle_velocity, coin_position)]))
