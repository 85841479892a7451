"""Keras-`Model`-shaped handles over the engine's networks.

The reference's BiGAN classes take factories `Callable[[int, int], Model]`
(src/bigan_basic.py:11-15) and only ever use `.predict`, `.summary`, `.layers[*].get_weights()`
and, in the tests, a few structural attributes (test/bigans_cc_test.py:16-55,89-121).  A
`NetModel` is that surface: it is created from a layer graph by the factories in
bigan_classify.py / bigan_cont.py and is bound to a `BiGanEngine` network once the BiGAN object
exists.  All arithmetic happens in the engine (CUDA); there is no host implementation here.
"""
import numpy as np


class _Activation:
    """Identity-comparable stand-ins for tf.nn.sigmoid / relu / softmax."""

    def __init__(self, name):
        self.__name__ = name

    def __repr__(self):
        return f"<activation {self.__name__}>"


class activations:
    sigmoid = _Activation("sigmoid")
    relu = _Activation("relu")
    softmax = _Activation("softmax")
    linear = _Activation("linear")

    @staticmethod
    def get(name):
        return getattr(activations, name if name != "none" else "linear")


class losses:
    class _Loss:
        def __init__(self, name):
            self.__name__ = name

        def __repr__(self):
            return f"<loss {self.__name__}>"

    binary_crossentropy = _Loss("binary_crossentropy")
    mse = _Loss("mse")


class RMSprop:
    """optimizers.RMSprop(learning_rate=0.0075, rho=0.85, momentum=0.1), src/bigan_classify.py:88.
    The update itself is cc_rmsprop_step (csrc/tail_kernels.cu); this object carries the
    hyper-parameters the five training graphs share."""

    def __init__(self, learning_rate=0.001, rho=0.9, momentum=0.0, epsilon=1e-7):
        self.learning_rate, self.rho, self.momentum, self.epsilon = learning_rate, rho, momentum, epsilon

    def get_config(self):
        return {"name": "RMSprop", "learning_rate": self.learning_rate, "rho": self.rho,
                "momentum": self.momentum, "epsilon": self.epsilon, "centered": False}


class InputLayer:
    def __init__(self, name, width):
        self.name = name
        self.input_shape = [(None, width)]
        self.output_shape = [(None, width)]

    def get_weights(self):
        return []


class LayerView:
    """One Dense / BatchNormalization layer of a bound network."""

    def __init__(self, model, index, kind, name, activation=None):
        self._model, self._index, self.kind, self.name = model, index, kind, name
        self.activation = activation

    def get_weights(self):
        net = self._model._net()
        net._wait_optimizer()
        net.gather_master()      # data parallel: collective, like NetModel.get_weights
        net._rows()
        L = net.layers[self._index]
        keys = ("w32", "b32") if L["kind"] == "dense" else (
            "gamma", "beta", "moving_mean", "moving_var")
        return [L[k].detach().float().cpu().numpy().copy() for k in keys]


class NetModel:
    """A network handle with the slice of the Keras Model API the reference uses."""

    def __init__(self, graph, role, input_names):
        self.graph = graph
        self.role = role                      # "G" | "E" | "D"
        self.name = graph.name
        self.trainable = True
        self._engine = None
        self._owner = None
        self._is_compiled = False
        self.loss = None
        self.optimizer = None
        g = graph
        self._inputs = [InputLayer(n, g.widths[g.inputs[k]]) for n, k in input_names]
        self.layers = list(self._inputs)
        di = bi = 0
        for i, lay in enumerate(g.layers):
            if lay[0] == "dense":
                self.layers.append(LayerView(self, i, "dense", f"dense_{di}",
                                             activations.get(lay[3])))
                di += 1
            else:
                self.layers.append(LayerView(self, i, "bn", f"batch_normalization_{bi}"))
                bi += 1

    # ---------------------------------------------------------------- structure
    @property
    def input_shape(self):
        s = [l.input_shape[0] for l in self._inputs]
        return s if len(s) > 1 else s[0]

    @property
    def output_shape(self):
        return (None, self.graph.widths[self.graph.output])

    def get_input_shape_at(self, _):
        return self.input_shape

    def get_output_shape_at(self, _):
        return self.output_shape

    def compile(self, optimizer=None, loss=None):
        self.optimizer, self.loss, self._is_compiled = optimizer, loss, True

    def count_params(self):
        n = 0
        for lay in self.graph.layers:
            n += (sum(lay[1]) * lay[2] + lay[2]) if lay[0] == "dense" else 4 * lay[1]
        return n

    def summary(self, print_fn=print):
        print_fn(f'Model: "{self.name}"')
        print_fn("_" * 65)
        print_fn(f"{'Layer (type)':<34}{'Output Shape':<20}{'Param #':>10}")
        print_fn("=" * 65)
        for l in self._inputs:
            print_fn(f"{l.name + ' (InputLayer)':<34}{str(l.input_shape):<20}{0:>10}")
        di = bi = 0
        for lay in self.graph.layers:
            if lay[0] == "dense":
                n = sum(lay[1]) * lay[2] + lay[2]
                print_fn(f"{f'dense_{di} (Dense, {lay[3]})':<34}{str((None, lay[2])):<20}{n:>10}")
                di += 1
            else:
                print_fn(f"{f'batch_normalization_{bi}':<34}{str((None, lay[1])):<20}{4 * lay[1]:>10}")
                bi += 1
        print_fn("=" * 65)
        print_fn(f"Total params: {self.count_params():,}")

    # ---------------------------------------------------------------- runtime
    def _bind(self, owner, engine):
        self._owner, self._engine = owner, engine

    def _net(self):
        if self._engine is None:
            raise RuntimeError(f"{self.name}: not bound to a BiGAN engine yet (the CUDA engine is "
                               f"created by ClassifyCellBiGan/ContinuousCellBiGan)")
        return self._engine.nets[self.role]

    def get_weights(self):
        return self._net().get_weights()

    def set_weights(self, arrays):
        self._net().set_weights(arrays)

    def weight_fingerprint(self):
        """Cheap device-side digest used by BasicBiGan.print_params_changes instead of the
        reference's deep copy of every layer (src/bigan_basic.py:72-81)."""
        self._net()._rows()      # one summation order whatever layout training left the state in
        p = self._net().p32
        return np.array([float(p.sum()), float((p * p).sum())])

    def train_on_batch(self, x, y=None):
        """Model.train_on_batch of a compiled component.  Only the discriminator is compiled on
        its own in the reference (src/bigan_classify.py:112-115); G and E are trained through
        the four combined graphs (`_train_gen_w_discr` ...)."""
        if self.role != "D" or not self._is_compiled:
            raise RuntimeError(f"{self.name} is not compiled for training on its own; use the "
                               f"BiGAN's combined training graphs")
        return self._owner._train_discriminator_on_batch(x, y)

    def predict(self, x, **_kwargs):
        """Model.predict: inference mode (BN moving stats, dropout off), float32 host result.
        Extra Keras kwargs (batch_size, use_multiprocessing, ...) are accepted and ignored."""
        return self._owner._predict(self.role, x)
