"""Run this WHERE TensorFlow 2.4 AND the reference checkout are available (they are not in
the build image) to pin the oracle against the real Keras implementation:

    PYTHONPATH=/path/to/cellcomm/src python tools/dump_keras_reference.py out.npz [Z G B]

It builds the reference's ContinuousCellBiGan, sets dropout rates to 0 (TF's dropout RNG
cannot be injected), dumps the initial weights in creation order, runs ONE trainings_step on
fixed inputs with fixed priors, and stores the three losses, the encodings and the post-step
weights.  tests can then load the .npz, set the same weights in oracle.OracleBiGan, run
trainings_step with all-ones masks and compare.  (File format: numpy .npz, keys g_w<i>,
e_w<i>, d_w<i>, x, z, r, losses, enc_after, g_after<i>, ...)
"""
import sys

import numpy as np


def creation_order(model):
    import re
    def key(layer):
        m = re.search(r"_(\d+)$", layer.name)
        return int(m.group(1)) if m else -1
    dense = sorted([l for l in model.layers if l.__class__.__name__ == "Dense"], key=key)
    bn = sorted([l for l in model.layers if l.__class__.__name__ == "BatchNormalization"], key=key)
    # interleave by the order the builder created them: both counters are global per process,
    # so sort all weighted layers by the order of their first variable's creation
    layers = [l for l in model.layers if l.weights]
    layers.sort(key=lambda l: l.weights[0]._unique_id if hasattr(l.weights[0], "_unique_id")
                else l.name)
    del dense, bn
    return layers


def main():
    import tensorflow as tf
    from bigan_cont import ContinuousCellBiGan
    out = sys.argv[1]
    Z, G, B = (int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (3, 400, 32)
    tf.random.set_seed(0)
    np.random.seed(0)
    net = ContinuousCellBiGan(Z, G)
    for comp in net.all_components:
        for l in comp.layers:
            if l.__class__.__name__ == "Dropout":
                l.rate = 0.0
    rng = np.random.default_rng(0)
    x = ((rng.random((B, G)) < 0.06) * (rng.poisson(1.2, (B, G)) + 1)).astype(np.float32)
    z = rng.random((B, Z), dtype=np.float32)
    r = rng.random((B, Z), dtype=np.float32)
    net.random_encoding_vector = lambda n: z
    net.random_uniform_vector = lambda n: r
    dump = {"x": x, "z": z, "r": r}
    for tag, comp in zip("ged", net.all_components):
        for i, l in enumerate(creation_order(comp)):
            for j, w in enumerate(l.get_weights()):
                dump[f"{tag}_w{i}_{j}"] = w
    losses = net.trainings_step(x)
    dump["losses"] = np.array([float(v) for v in losses])
    dump["enc_after"] = net.encoding_prediction(x)
    for tag, comp in zip("ged", net.all_components):
        for i, l in enumerate(creation_order(comp)):
            for j, w in enumerate(l.get_weights()):
                dump[f"{tag}_after{i}_{j}"] = w
    np.savez_compressed(out, **dump)
    print("wrote", out, "losses", dump["losses"])


if __name__ == "__main__":
    main()
