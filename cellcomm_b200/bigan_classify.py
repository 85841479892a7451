"""Drop-in for the reference's `src/bigan_classify.py`: the categorical-code BiGAN, the
network factories (`_build_generator/_build_encoder/_build_discriminator`, reference :10-75),
the five compiled training graphs with their freeze pattern (:83-115) and `trainings_step`
(:126-155).

The factories return `NetModel` handles over layer graphs (cellcomm_b200.engine); the
arithmetic of `train_on_batch` / `predict` runs in libcellcomm_b200.so through
`BiGanEngine` (tcgen05 GEMMs + tail kernels).  If a factory returns something else (the
reference's tests inject mocks), no engine is built and only the wiring is available.
"""
import os
from typing import Callable

import numpy as np

try:
    from .bigan_basic import BasicBiGan
    from . import engine as _engine
    from .models import NetModel, RMSprop, losses, activations  # noqa: F401
    from .cell_type_training import CellBatch, CellMatrix
except ImportError:  # imported as top-level modules (PYTHONPATH=src style)
    from bigan_basic import BasicBiGan
    from cellcomm_b200 import engine as _engine
    from cellcomm_b200.models import NetModel, RMSprop, losses, activations  # noqa: F401
    from cellcomm_b200.cell_type_training import CellBatch, CellMatrix


def _build_generator(encoding_size, gene_size):
    return NetModel(_engine.classify_generator_graph(encoding_size, gene_size), "G",
                    [("gen_encoding_in", "z"), ("gen_random_in", "r")])


def _build_encoder(encoding_size, gene_size):
    return NetModel(_engine.classify_encoder_graph(encoding_size, gene_size), "E",
                    [("enc_cell_in", "cell")])


def _build_discriminator(encoding_size, gene_size):
    return NetModel(_engine.discriminator_graph(encoding_size, gene_size), "D",
                    [("encoding_input", "z"), ("cell_input", "cell")])


def print_dot():
    print('.', end='', flush=True)


class TrainGraph:
    """One of the reference's five compiled training models: which component it updates, which
    it runs frozen, its loss and the (shared) optimizer.  `train_on_batch` runs the matching
    sub-step of the engine."""

    def __init__(self, owner, name, substep, trained, layers, input_shape, loss, optimizer):
        self._owner, self.name, self._substep, self._trained = owner, name, substep, trained
        self.layers, self.input_shape, self.output_shape = layers, input_shape, layers[-1].output_shape
        self.loss, self.optimizer, self._is_compiled = loss, optimizer, True

    def train_on_batch(self, x, y=None):
        return self._owner._train_on_batch(self._substep, x, y)


class ClassifyCellBiGan(BasicBiGan):
    VARIANT = "classify"

    def __init__(self, encoding_size, gene_size,
                 generator_factory: Callable[[int, int], object] = _build_generator,
                 encoder_factory: Callable[[int, int], object] = _build_encoder,
                 discriminator_factory: Callable[[int, int], object] = _build_discriminator,
                 device=None, seed=None, dist=None):
        super().__init__(encoding_size, gene_size, generator_factory, encoder_factory, discriminator_factory)
        self.gene_size = gene_size
        discr_optimizer = RMSprop(learning_rate=0.0075, rho=0.85, momentum=0.1)
        self._optimizer = discr_optimizer
        G, E, D = self._generator, self._encoder, self._discriminator
        self._engine = None
        self._staged = None
        if all(isinstance(m, NetModel) for m in (G, E, D)):
            if device is None:
                device = os.environ.get("CELLCOMM_B200_DEVICE", "cuda")
            if dist is None:
                dist = _engine.default_dist()
            self._engine = _engine.BiGanEngine(
                self.VARIANT, encoding_size, gene_size, max_batch=1, device=device, seed=seed,
                dist=dist, graphs={"G": G.graph, "E": E.graph, "D": D.graph})
            for m in (G, E, D):
                m._bind(self, self._engine)
            z = (None, encoding_size)
            bce, mse = losses.binary_crossentropy, losses.mse
            # freeze pattern of the five graphs, reference :90-115
            self._train_gen_w_discr = TrainGraph(self, 'train-generator-with-discriminator', 1, "G",
                                                 [G, D], [z, z], bce, discr_optimizer)
            self._train_gen_w_enc = TrainGraph(self, 'train-generator-with-encoder', 2, "G",
                                               [E, G], [(None, gene_size), z], mse, discr_optimizer)
            self._train_enc_w_discr = TrainGraph(self, 'train-encoder-with-discriminator', 3, "E",
                                                 [E, D], (None, gene_size), bce, discr_optimizer)
            self._train_enc_w_gen = TrainGraph(self, 'train-encoder-with-generator', 4, "E",
                                               [G, E], [z, z], mse, discr_optimizer)
            D.compile(optimizer=discr_optimizer, loss=bce)
            G.trainable, E.trainable, D.trainable = False, False, True

    # ------------------------------------------------------------------ priors
    def random_encoding_vector(self, batch_size):
        """to_categorical(np.random.randint(0, Z, B), Z) on numpy's GLOBAL state (reference
        :117-119; golden np.random.seed(21) -> [1,3,0,0,0], test/bigans_cc_test.py:57-63)."""
        rand_ixs = np.random.randint(0, self.encoding_size, batch_size)
        out = np.zeros((batch_size, self.encoding_size), dtype=np.float32)
        out[np.arange(batch_size), rand_ixs] = 1.0
        return out

    def trainings_encoding_prediction(self, cell_data):
        prediction = np.asarray(self.encoding_prediction(cell_data))
        argmax = np.argmax(prediction, -1)
        out = np.zeros((len(argmax), self.encoding_size), dtype=np.float32)
        out[np.arange(len(argmax)), argmax] = 1.0
        return out

    # ------------------------------------------------------------------ the step
    def trainings_step(self, batch):
        """Six updates + two predicts, reference :126-142, as one pass of BiGanEngine.train_step
        (the sub-step order, freeze pattern and loss bookkeeping live in engine.py)."""
        eng = self._require_engine()
        batch_size = len(batch)
        encodings = self.random_encoding_vector(batch_size)
        noise = self.random_uniform_vector(batch_size)
        if (isinstance(batch, CellBatch) and eng.device.type == "cuda" and
                eng.peer_graphable() and os.environ.get("CELLCOMM_B200_GRAPH", "1") != "0"):
            # single GPU: the whole step (gather + ~850 kernels) is one CUDA-graph launch
            gs = eng.capture_step(batch.matrix.device_csr(eng.device), batch.matrix.shape[1],
                                  batch_size, latents="host")
            eng.z32[:batch_size].copy_(_as_f32(encodings), non_blocking=True)
            eng.r32[:batch_size].copy_(_as_f32(noise), non_blocking=True)
            return gs.replay(batch.positions)
        x16 = self._stage_cells(batch)
        eng.set_latents(encodings, noise, batch_size)
        g_loss, e_loss, d_loss = eng.train_step(x16)
        return g_loss, e_loss, d_loss

    # ------------------------------------------------------------------ checkpoint / resume
    def save_checkpoint(self, path):
        """Weights, RMSprop slots, BN moving statistics and RNG position -> one .npz (the
        reference has no checkpointing)."""
        self._require_engine().save_checkpoint(path)

    def load_checkpoint(self, path):
        self._require_engine().load_checkpoint(path)

    # ------------------------------------------------------------------ plumbing
    def _require_engine(self):
        if self._engine is None:
            raise RuntimeError("this BiGAN was built from non-NetModel components (mocks): the "
                               "CUDA engine is unavailable and there is no host fallback")
        return self._engine

    def _stage_cells(self, cells, row_start=0, rows=None):
        """[n, gene_size] bf16 device tile for `cells`: a CellBatch / CellMatrix is gathered
        from the device-resident CSR (cc_gather_rows); anything array-like (DataFrame, ndarray,
        list, torch tensor) is uploaded as float32 and cast on the device."""
        import torch
        ops = _engine.ops
        eng = self._engine
        dev = eng.device
        if isinstance(cells, (CellBatch, CellMatrix)):
            mat = cells.matrix if isinstance(cells, CellBatch) else cells
            rowptr, colidx, values = mat.device_csr(dev)
            if isinstance(cells, CellBatch):
                n = len(cells)
                idx = torch.from_numpy(cells.positions).to(dev, non_blocking=True)
                out = self._tile(n)
                ops.gather_rows(rowptr, colidx, values, mat.shape[1], row_idx=idx, out16=out)
            else:
                n = len(mat) - row_start if rows is None else rows
                out = self._tile(n)
                ops.gather_rows(rowptr, colidx, values, mat.shape[1], row_start=row_start,
                                n_rows=n, out16=out)
            return out
        if isinstance(cells, torch.Tensor):
            t = cells
        else:
            arr = cells.values if hasattr(cells, "values") and not isinstance(cells, np.ndarray) \
                else cells
            t = torch.as_tensor(np.ascontiguousarray(np.asarray(arr, dtype=np.float32)))
        if rows is not None:
            t = t[row_start:row_start + rows]
        t = t.to(dev, dtype=torch.float32, non_blocking=True)
        if t.dim() != 2 or t.shape[1] != self.gene_size:
            raise ValueError(f"expected cells of shape (n, {self.gene_size}), got {tuple(t.shape)}")
        out = self._tile(t.shape[0])
        ops.cast_f32_to_bf16(t, out)
        return out

    def _tile(self, n):
        import torch
        buf = getattr(self, "_tile_buf", None)
        if buf is None or buf.shape[0] < n:
            self._tile_buf = buf = _engine.ops.alloc2d(max(n, 1), self.gene_size,
                                                       device=self._engine.device)
        return buf[:n]

    PREDICT_TILE = 4096

    def _predict(self, role, x):
        """Model.predict for G / E / D in inference mode -> float32 host array."""
        import torch
        eng = self._require_engine()
        dev = eng.device
        if role == "E":
            n = len(x)
            out = torch.empty((n, self.encoding_size), dtype=torch.float32, device=dev)
            for s in range(0, n, self.PREDICT_TILE):
                m = min(self.PREDICT_TILE, n - s)
                tile = self._stage_cells(x[s:s + m] if _sliceable(x) else x, s, m) \
                    if not isinstance(x, CellBatch) else self._stage_cells(
                        CellBatch(x.matrix, x.positions[s:s + m]))
                eng.encode(tile, out32=out[s:s + m])
            return out.cpu().numpy()
        enc, second = x
        z = torch.as_tensor(np.asarray(enc, dtype=np.float32)).to(dev)
        n = z.shape[0]
        if role == "G":
            r = torch.as_tensor(np.asarray(second, dtype=np.float32)).to(dev)
            out = torch.empty((n, self.gene_size), dtype=torch.float32, device=dev)
            for s in range(0, n, self.PREDICT_TILE):
                m = min(self.PREDICT_TILE, n - s)
                eng.set_latents(z[s:s + m], r[s:s + m], m)
                eng.generate(m, out32=out[s:s + m])
            return out.cpu().numpy()
        out = torch.empty((n, 1), dtype=torch.float32, device=dev)
        for s in range(0, n, self.PREDICT_TILE):
            m = min(self.PREDICT_TILE, n - s)
            tile = self._stage_cells(second[s:s + m] if _sliceable(second) else second, s, m) \
                if not isinstance(second, CellBatch) else self._stage_cells(
                    CellBatch(second.matrix, second.positions[s:s + m]))
            eng.discriminate(z[s:s + m].contiguous(), tile, out[s:s + m])
        return out.cpu().numpy()

    def _train_on_batch(self, substep, x, y):
        """Model.train_on_batch of one of the compiled graphs: runs that single sub-step and
        returns its loss (the target `y` is implied by the graph, as in the reference's calls
        :144-155)."""
        import torch
        eng = self._require_engine()
        ops = _engine.ops
        if substep in (1, 4):
            enc, noise = x
            n = len(enc)
            eng.set_latents(enc, noise, n)
            cells = self._tile(n)
        elif substep == 2:
            cells_in, noise = x
            n = len(cells_in)
            eng.set_latents(np.zeros((n, self.encoding_size), np.float32), noise, n)
            cells = self._stage_cells(cells_in)
        else:
            n = len(x)
            eng.reserve(n)
            cells = self._stage_cells(x)
        ops.fill_f32(eng.loss_buf, 0.0)
        eng.substep(substep, cells)
        slot = {1: 0, 2: 1, 3: 2, 4: 3, 6: 4, 8: 5}[substep]
        return float(eng.loss_buf[slot])


def _as_f32(a):
    import torch
    return torch.as_tensor(np.asarray(a, dtype=np.float32))


def _sliceable(x):
    return not isinstance(x, (CellMatrix,))
