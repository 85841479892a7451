"""ORACLE — test infrastructure, NOT product code.

CPU restatement (torch-CPU, fp32 or fp64) of the reference's BiGAN hot path:

  * networks        src/bigan_cont.py:7-41 (Continuous G/E), src/bigan_classify.py:10-75
                    (Classify G/E and the discriminator D shared by both variants)
  * five compiled training graphs + freeze pattern   src/bigan_classify.py:83-115
  * trainings_step (six updates + two predicts)      src/bigan_classify.py:126-155
  * encoding_prediction / generate_cells / accuracy  src/bigan_basic.py:29-64
  * priors                                           src/bigan_basic.py:36-37,
                                                     src/bigan_cont.py:52-53,
                                                     src/bigan_classify.py:117-119

The arithmetic itself lives in the third-party dependency tensorflow==2.4.0
(/root/reference/requirements.txt:3), which is NOT vendored and NOT installable here.  This
file restates the published Keras 2.4.0 semantics those call sites imply:
Dense = act(xW+b) with glorot_uniform/zeros init; BatchNormalization(momentum=0.99, eps=1e-3,
biased batch variance, frozen => inference mode); Dropout active inside frozen sub-models
during train_on_batch and off in predict; binary_crossentropy on a Sigmoid output = sigmoid
cross-entropy with logits, mean over the batch; mse = mean over all elements; RMSprop with
momentum (eps inside the sqrt, lr inside the momentum buffer, zero-initialised slots).

PARITY STATUS: **parity unpinned for the NN arithmetic** — the reference's tests hold no
golden loss / gradient / weight / encoding, and TensorFlow cannot be run
here.  What IS pinned by the reference's own tests (and checked in tests/test_oracle_goldens.py):
round-half-even golden, np.random.seed(21) one-hot golden, one-hot-argmax golden, prior range,
accuracy bookkeeping, loader/sampler goldens (oracle/loader_oracle.py).  Gradients are checked
against fp64 finite differences.  tools/dump_keras_reference.py lets anyone with TF 2.4 pin it.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  Every random quantity is an explicit input: batch rows, `encodings`,
`noise`, and one keep-mask per Dropout call per sub-step.
"""
import math

import numpy as np
import torch

LR, RHO, MOMENTUM, EPSILON = 0.0075, 0.85, 0.1, 1e-7   # src/bigan_classify.py:88 + Keras defaults
BN_MOMENTUM, BN_EPS = 0.99, 1e-3                        # Keras BatchNormalization defaults
REAL_LABEL = 0.95                                       # src/bigan_classify.py:128


# --------------------------------------------------------------------------- specs
# Layer lists in CREATION ORDER inside each reference builder.  ("dense", in, out, act) and
# ("bn", width); dropout / concat structure lives in the forward functions below.
def cont_generator_spec(Z, G):
    w = [int(G * f) for f in (0.2, 0.1)]                # src/bigan_cont.py:8
    return [("dense", 2 * Z, 50, "sigmoid"), ("dense", 50 + 2 * Z, 256, "sigmoid"),
            ("dense", 256, 256, "sigmoid"), ("bn", 256),
            ("dense", 256 + 2 * Z, w[1], "sigmoid"), ("dense", w[1], w[0], "relu"),
            ("bn", w[0]), ("dense", w[0], G, "relu")]


def cont_encoder_spec(Z, G):
    w = [int(G * f) for f in (0.1, 0.05)]               # src/bigan_cont.py:29
    return [("dense", G, w[0], "sigmoid"), ("dense", w[0] + G, w[1], "sigmoid"), ("bn", w[1]),
            ("dense", w[1], 150, "sigmoid"), ("dense", 150, 150, "sigmoid"),
            ("dense", 150, Z, "sigmoid")]


def classify_generator_spec(Z, G):                      # src/bigan_classify.py:10-25
    return [("dense", 2 * Z, 50, "sigmoid"), ("dense", 50 + 2 * Z, 256, "sigmoid"), ("bn", 256),
            ("dense", 256 + 2 * Z, 256, "sigmoid"), ("dense", 256, 1024, "relu"),
            ("dense", 1024, G, "relu")]


def classify_encoder_spec(Z, G):                        # src/bigan_classify.py:28-40
    return [("dense", G, 1000, "sigmoid"), ("dense", 1000, 300, "sigmoid"),
            ("dense", 1300, 150, "sigmoid"), ("dense", 150, Z, "softmax")]


def discriminator_spec(Z, G):                           # src/bigan_classify.py:43-75
    l = [int(G * f) for f in (0.3, 0.1, 0.05)]
    return [("dense", Z, 50, "sigmoid"), ("dense", Z, 50, "sigmoid"), ("bn", 100 + Z),
            ("dense", 100 + Z, 256, "sigmoid"), ("dense", 256, 256, "sigmoid"),
            ("dense", 256, 256, "sigmoid"),
            ("dense", G, l[0], "sigmoid"), ("dense", l[0] + G, l[1], "sigmoid"), ("bn", l[1]),
            ("dense", l[1], l[2], "sigmoid"), ("dense", l[2], 256, "sigmoid"),
            ("dense", 256, 256, "sigmoid"),
            ("dense", 512, 300, "sigmoid"), ("bn", 300), ("dense", 300, 50, "sigmoid"),
            ("dense", 50, 50, "sigmoid"), ("dense", 50, 10, "sigmoid"), ("dense", 10, 1, "sigmoid")]


DROPOUT_RATES = {   # per net, in call order
    ("cont", "G"): [0.1, 0.1], ("cont", "E"): [0.15, 0.1],
    ("classify", "G"): [0.1, 0.1, 0.1], ("classify", "E"): [0.15, 0.15],
    "D": [0.15, 0.15, 0.15, 0.15],
}


def dropout_widths(variant, net, Z, G):
    """Width of the tensor each Dropout call acts on (for building injected masks)."""
    if net == "D":
        l = [int(G * f) for f in (0.3, 0.1, 0.05)]
        return [100 + Z, l[0], l[1], 300]
    if variant == "cont":
        if net == "G":
            return [50 + 2 * Z, 256 + 2 * Z]
        w = [int(G * f) for f in (0.1, 0.05)]
        return [w[0], w[1]]
    if net == "G":
        return [50 + 2 * Z, 256 + 2 * Z, 256]
    return [1000, 300]


def init_net(spec, gen, dtype=torch.float32):
    """Keras initialisers: Dense glorot_uniform kernel + zero bias; BN gamma=1, beta=0,
    moving_mean=0, moving_variance=1.  Returns a list of per-layer dicts."""
    layers = []
    for item in spec:
        if item[0] == "dense":
            _, fi, fo, act = item
            limit = math.sqrt(6.0 / (fi + fo)) if fi + fo > 0 else 0.0
            k = (torch.rand(fi, fo, generator=gen, dtype=torch.float64) * 2 - 1) * limit
            layers.append({"kind": "dense", "act": act, "kernel": k.to(dtype),
                           "bias": torch.zeros(fo, dtype=dtype)})
        else:
            n = item[1]
            layers.append({"kind": "bn", "gamma": torch.ones(n, dtype=dtype),
                           "beta": torch.zeros(n, dtype=dtype),
                           "moving_mean": torch.zeros(n, dtype=dtype),
                           "moving_var": torch.ones(n, dtype=dtype)})
    return layers


def trainable_params(layers):
    out = []
    for l in layers:
        out += [l["kernel"], l["bias"]] if l["kind"] == "dense" else [l["gamma"], l["beta"]]
    return out


# --------------------------------------------------------------------------- layer math
def _act(x, name):
    if name == "sigmoid":
        return torch.sigmoid(x)
    if name == "relu":
        return torch.relu(x)
    if name == "softmax":
        return torch.softmax(x, -1)
    return x


class _Ctx:
    """Per-call context: BN mode, dropout masks (None => predict mode, dropout off), and the
    list of BN moving-stat updates to commit after the step."""

    def __init__(self, bn_train, masks, n_total=None):
        self.bn_train = bn_train
        self.masks = list(masks) if masks is not None else None
        self.mask_i = 0
        self.bn_updates = []

    def dense(self, layer, x, pre_activation=False):
        z = x @ layer["kernel"] + layer["bias"]
        return z if pre_activation else _act(z, layer["act"])

    def dropout(self, x, rate):
        if self.masks is None:
            return x
        m = self.masks[self.mask_i]
        self.mask_i += 1
        assert m.shape == x.shape, f"dropout mask shape {tuple(m.shape)} != {tuple(x.shape)}"
        return x * m.to(x.dtype) / (1.0 - rate)

    def bn(self, layer, x):
        if self.bn_train:
            mean = x.mean(0)
            var = ((x - mean) ** 2).mean(0)
            self.bn_updates.append((layer, mean.detach(), var.detach()))
        else:
            mean, var = layer["moving_mean"], layer["moving_var"]
        return (x - mean) / torch.sqrt(var + BN_EPS) * layer["gamma"] + layer["beta"]

    def commit_bn(self):
        for layer, mean, var in self.bn_updates:
            layer["moving_mean"].mul_(BN_MOMENTUM).add_(mean * (1 - BN_MOMENTUM))
            layer["moving_var"].mul_(BN_MOMENTUM).add_(var * (1 - BN_MOMENTUM))


def cont_generator(L, ctx, z, r):                       # src/bigan_cont.py:7-25
    r0, r1 = DROPOUT_RATES[("cont", "G")]
    all_in = torch.cat([z, r], 1)
    x = ctx.dense(L[0], all_in)
    x = ctx.dropout(torch.cat([x, all_in], 1), r0)
    x = ctx.dense(L[1], x)
    x = ctx.dense(L[2], x)
    x = ctx.bn(L[3], x)
    x = ctx.dropout(torch.cat([x, all_in], 1), r1)
    x = ctx.dense(L[4], x)
    x = ctx.dense(L[5], x)
    x = ctx.bn(L[6], x)
    return ctx.dense(L[7], x)


def cont_encoder(L, ctx, cell):                         # src/bigan_cont.py:28-41
    r0, r1 = DROPOUT_RATES[("cont", "E")]
    x = ctx.dense(L[0], cell)
    x = ctx.dropout(x, r0)
    x = ctx.dense(L[1], torch.cat([x, cell], 1))
    x = ctx.dropout(x, r1)
    x = ctx.bn(L[2], x)
    x = ctx.dense(L[3], x)
    x = ctx.dense(L[4], x)
    return ctx.dense(L[5], x)


def classify_generator(L, ctx, z, r):                   # src/bigan_classify.py:10-25
    r0, r1, r2 = DROPOUT_RATES[("classify", "G")]
    all_in = torch.cat([z, r], 1)
    x = ctx.dense(L[0], all_in)
    x = ctx.dropout(torch.cat([x, all_in], 1), r0)
    x = ctx.dense(L[1], x)
    x = ctx.bn(L[2], x)
    x = ctx.dropout(torch.cat([x, all_in], 1), r1)
    x = ctx.dense(L[3], x)
    x = ctx.dropout(x, r2)
    x = ctx.dense(L[4], x)
    return ctx.dense(L[5], x)


def classify_encoder(L, ctx, cell):                     # src/bigan_classify.py:28-40
    r0, r1 = DROPOUT_RATES[("classify", "E")]
    proc = ctx.dense(L[0], cell)
    x = ctx.dropout(proc, r0)
    x = ctx.dense(L[1], x)
    x = ctx.dropout(x, r1)
    x = ctx.dense(L[2], torch.cat([x, proc], 1))
    return ctx.dense(L[3], x)


def discriminator(L, ctx, z, cell, logits=False):       # src/bigan_classify.py:43-75
    r = DROPOUT_RATES["D"]
    a = ctx.dense(L[0], z)
    b = ctx.dense(L[1], z)
    x = ctx.dropout(torch.cat([a, b, z], 1), r[0])
    x = ctx.bn(L[2], x)
    x = ctx.dense(L[3], x)
    x = ctx.dense(L[4], x)
    se = ctx.dense(L[5], x)
    x = ctx.dense(L[6], cell)
    x = ctx.dropout(x, r[1])
    x = ctx.dense(L[7], torch.cat([x, cell], 1))
    x = ctx.bn(L[8], x)
    x = ctx.dropout(x, r[2])
    x = ctx.dense(L[9], x)
    x = ctx.dense(L[10], x)
    sg = ctx.dense(L[11], x)
    x = ctx.dense(L[12], torch.cat([se, sg], 1))
    x = ctx.bn(L[13], x)
    x = ctx.dropout(x, r[3])
    x = ctx.dense(L[14], x)
    x = ctx.dense(L[15], x)
    x = ctx.dense(L[16], x)
    return ctx.dense(L[17], x, pre_activation=logits)


# --------------------------------------------------------------------------- losses / optimiser
def bce_from_logits(logits, target):
    """losses.binary_crossentropy on a Sigmoid output (Keras backend takes the op's logits):
    mean over the last axis, then over the batch."""
    t = torch.full_like(logits, target)
    per = torch.clamp(logits, min=0) - logits * t + torch.log1p(torch.exp(-logits.abs()))
    return per.mean(-1).mean()


def mse(pred, target):
    return ((pred - target) ** 2).mean(-1).mean()


def round_half_even(x):                                 # tf.math.round, src/bigan_basic.py:44
    return torch.round(x)


def to_categorical_argmax(p):                           # src/bigan_classify.py:121-124
    idx = torch.argmax(p, -1)
    return torch.nn.functional.one_hot(idx, p.shape[-1]).to(p.dtype)


def rmsprop_apply(params, grads, slots):
    """Keras RMSprop(momentum>0), non-centered:  ms = rho*ms + (1-rho) g^2 ;
    mom = momentum*mom + lr*g/sqrt(ms+eps) ; w -= mom.  Slots are per variable, zero-init."""
    for p, g in zip(params, grads):
        key = id(p)
        if key not in slots:
            slots[key] = (torch.zeros_like(p), torch.zeros_like(p))
        ms, mom = slots[key]
        ms.mul_(RHO).add_((1 - RHO) * g * g)
        mom.mul_(MOMENTUM).add_(LR * g / torch.sqrt(ms + EPSILON))
        p.sub_(mom)


# --------------------------------------------------------------------------- the model
class OracleBiGan:
    """variant: 'cont' (ContinuousCellBiGan) or 'classify' (ClassifyCellBiGan)."""

    def __init__(self, variant, encoding_size, gene_size, seed=0, dtype=torch.float32):
        assert variant in ("cont", "classify")
        self.variant, self.Z, self.G, self.dtype = variant, encoding_size, gene_size, dtype
        gen = torch.Generator().manual_seed(seed)
        gspec = cont_generator_spec if variant == "cont" else classify_generator_spec
        espec = cont_encoder_spec if variant == "cont" else classify_encoder_spec
        self.gen_layers = init_net(gspec(self.Z, self.G), gen, dtype)
        self.enc_layers = init_net(espec(self.Z, self.G), gen, dtype)
        self.dis_layers = init_net(discriminator_spec(self.Z, self.G), gen, dtype)
        self.slots = {}
        self.last_grads = {}

    # -- plumbing
    def nets(self):
        return {"G": self.gen_layers, "E": self.enc_layers, "D": self.dis_layers}

    def _G(self, ctx, z, r):
        f = cont_generator if self.variant == "cont" else classify_generator
        return f(self.gen_layers, ctx, z, r)

    def _E(self, ctx, x):
        f = cont_encoder if self.variant == "cont" else classify_encoder
        return f(self.enc_layers, ctx, x)

    def _D(self, ctx, z, x, logits=False):
        return discriminator(self.dis_layers, ctx, z, x, logits)

    def t(self, a):
        if isinstance(a, torch.Tensor):
            return a.to(self.dtype)
        return torch.as_tensor(np.asarray(a), dtype=self.dtype)

    # -- inference API (src/bigan_basic.py:29-64)
    def encoding_prediction(self, cells):
        with torch.no_grad():
            return self._E(_Ctx(False, None), self.t(cells))

    def generator_predict(self, z, r):
        with torch.no_grad():
            return self._G(_Ctx(False, None), self.t(z), self.t(r))

    def generate_cells(self, z, r):
        return round_half_even(self.generator_predict(z, r))

    def discriminator_predict(self, z, x):
        with torch.no_grad():
            return self._D(_Ctx(False, None), self.t(z), self.t(x))

    def trainings_encoding_prediction(self, cells):
        p = self.encoding_prediction(cells)
        return p if self.variant == "cont" else to_categorical_argmax(p)

    def evaluate_discriminator_accuracy(self, batch, encodings, noise):
        """(true positives, true negatives), src/bigan_basic.py:50-64."""
        gen = self.generate_cells(encodings, noise)
        fake = self.discriminator_predict(encodings, gen)
        false_neg = int(torch.count_nonzero(torch.round(fake)))
        real = self.discriminator_predict(self.encoding_prediction(batch), batch)
        return int(torch.count_nonzero(torch.round(real))), len(batch) - false_neg

    # -- one compiled-graph update
    def _update(self, name, trained, loss_fn, masks):
        """Run one train_on_batch: forward with the sub-step's modes, backward, RMSprop on the
        trained net, BN moving-stat update for the trained net only."""
        layers = self.nets()[trained]
        params = trainable_params(layers)
        for p in params:
            p.requires_grad_(True)
        ctxs = {}

        def ctx_for(net):
            if net not in ctxs:
                ctxs[net] = _Ctx(net == trained, masks.get(net) if masks is not None else [])
            return ctxs[net]

        loss = loss_fn(ctx_for)
        grads = torch.autograd.grad(loss, params, allow_unused=True)
        grads = [torch.zeros_like(p) if g is None else g for p, g in zip(params, grads)]
        for p in params:
            p.requires_grad_(False)
        with torch.no_grad():
            self.last_grads[name] = [g.clone() for g in grads]
            rmsprop_apply(params, grads, self.slots)
            ctxs[trained].commit_bn()
        return float(loss.detach())

    def substep(self, k, x, z, r, masks=None):
        """Sub-step k (1..8) of trainings_step; returns the loss for the six updates."""
        B = x.shape[0]
        mk = lambda s: self._masks_for(s, B, masks)
        if k == 1:   # _train_gen_w_discr: BCE(D(z, G(z,r)), 0.95), updates G          :145
            return self._update("1", "G", lambda c: bce_from_logits(
                self._D(c("D"), z, self._G(c("G"), z, r), logits=True), REAL_LABEL), mk(1))
        if k == 2:   # _train_gen_w_enc: MSE(G(E(x), r), x), updates G                  :146
            return self._update("2", "G", lambda c: mse(
                self._G(c("G"), self._E(c("E"), x), r), x), mk(2))
        if k == 3:   # _train_enc_w_discr: BCE(D(E(x), x), 0), updates E                :150
            return self._update("3", "E", lambda c: bce_from_logits(
                self._D(c("D"), self._E(c("E"), x), x, logits=True), 0.0), mk(3))
        if k == 4:   # _train_enc_w_gen: MSE(E(G(z,r)), z), updates E                   :151
            return self._update("4", "E", lambda c: mse(
                self._E(c("E"), self._G(c("G"), z, r)), z), mk(4))
        if k == 5:   # generated_cells = round(G.predict((z, r)))                       :136
            self.gen_cells = self.generate_cells(z, r)
            return None
        if k == 6:   # D.train_on_batch((z, generated_cells), zeros)                    :137
            return self._update("6", "D", lambda c: bce_from_logits(
                self._D(c("D"), z, self.gen_cells, logits=True), 0.0), mk(6))
        if k == 7:   # generated_encodings = trainings_encoding_prediction(batch)       :138
            self.gen_enc = self.trainings_encoding_prediction(x)
            return None
        if k == 8:   # D.train_on_batch((generated_encodings, batch), 0.95)             :139
            return self._update("8", "D", lambda c: bce_from_logits(
                self._D(c("D"), self.gen_enc, x, logits=True), REAL_LABEL), mk(8))
        raise ValueError(k)

    def trainings_step(self, batch, encodings, noise, masks=None):
        """src/bigan_classify.py:126-155.  masks: {substep: {net: [keep masks in call order]}}
        for substeps 1,2,3,4,6,8; missing entries => all-ones masks (nothing dropped, the
        1/(1-rate) rescale still applies)."""
        x = self.t(batch)
        z = self.t(encodings)
        r = self.t(noise)
        l = {k: self.substep(k, x, z, r, masks) for k in (1, 2, 3, 4, 5, 6, 7, 8)}
        self.last_losses = {str(k): l[k] for k in (1, 2, 3, 4, 6, 8)}
        self.last_aux = {"generated_cells": self.gen_cells, "generated_encodings": self.gen_enc}
        return l[1] + l[2], l[3] + l[4], float(np.mean([l[6], l[8]]))

    def _masks_for(self, substep, B, masks):
        nets = {1: ("G", "D"), 2: ("E", "G"), 3: ("E", "D"), 4: ("G", "E"), 6: ("D",), 8: ("D",)}
        out = {}
        given = (masks or {}).get(substep, {})
        for net in nets[substep]:
            if net in given:
                out[net] = [self.t(m) for m in given[net]]
            else:
                out[net] = [torch.ones(B, w, dtype=self.dtype)
                            for w in dropout_widths(self.variant, net, self.Z, self.G)]
        return out

    # -- weights in creation order: Dense -> [kernel, bias]; BN -> [gamma, beta, mean, var]
    def get_weights(self, net):
        out = []
        for l in self.nets()[net]:
            keys = ("kernel", "bias") if l["kind"] == "dense" else (
                "gamma", "beta", "moving_mean", "moving_var")
            out += [l[k].detach().clone() for k in keys]
        return out

    def get_slots(self, net):
        """[(ms, mom)] per trainable tensor, creation order (zeros before the first update)."""
        out = []
        for p in trainable_params(self.nets()[net]):
            ms, mom = self.slots.get(id(p), (torch.zeros_like(p), torch.zeros_like(p)))
            out.append((ms.clone(), mom.clone()))
        return out

    def set_slots(self, net, slots):
        """[(ms, mom)] per trainable tensor, creation order (the inverse of get_slots)."""
        for p, (ms, mom) in zip(trainable_params(self.nets()[net]), slots):
            self.slots[id(p)] = (self.t(ms).clone(), self.t(mom).clone())

    def set_weights(self, net, arrays):
        it = iter(arrays)
        for l in self.nets()[net]:
            keys = ("kernel", "bias") if l["kind"] == "dense" else (
                "gamma", "beta", "moving_mean", "moving_var")
            for k in keys:
                l[k] = self.t(next(it)).clone()
        self.slots = {}


# priors (host RNG)  ------------------------------------------------------------------------
def classify_random_encoding_vector(encoding_size, batch_size):
    """to_categorical(np.random.randint(0, Z, B), Z) on the GLOBAL numpy state,
    src/bigan_classify.py:117-119."""
    idx = np.random.randint(0, encoding_size, batch_size)
    out = np.zeros((batch_size, encoding_size), dtype=np.float32)
    out[np.arange(batch_size), idx] = 1.0
    return out


def make_masks(variant, Z, G, B, seed):
    """Bernoulli keep-masks for every Dropout call of one trainings_step (28 for 'cont')."""
    g = torch.Generator().manual_seed(seed)
    nets = {1: ("G", "D"), 2: ("E", "G"), 3: ("E", "D"), 4: ("G", "E"), 6: ("D",), 8: ("D",)}
    out = {}
    for s, ns in nets.items():
        out[s] = {}
        for net in ns:
            rates = DROPOUT_RATES["D"] if net == "D" else DROPOUT_RATES[(variant, net)]
            widths = dropout_widths(variant, net, Z, G)
            out[s][net] = [(torch.rand(B, w, generator=g) >= rate).to(torch.uint8)
                           for w, rate in zip(widths, rates)]
    return out
