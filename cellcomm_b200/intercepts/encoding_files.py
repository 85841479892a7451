"""Per-iteration encoding files `logs/<run-id>/encodings/<iteration>.enc`.

The reference's offline tools read these files -- `import_encodings.py:21-27,66-77` (Mongo
importer) and `convert_encodings_to_mp4.py:27-30` (plot/movie maker): each is
`pickle.dump` of the float32 `(n_cells, encoding_size)` array that `encoding_prediction`
returns for ALL cells, un-scaled (the readers multiply by 255 themselves).  Nothing in the
reference's `src/` writes them any more; this interceptor is the missing writer,
fed by the same encode-all-cells pass as `DbRecorder.intercept` (db_recorder.py:85).
"""
import os
import pickle

import numpy as np


class EncodingFiles:
    def __init__(self, log_dir):
        """log_dir = logs/<run-id> (the directory `SinkIntercepts` writes losses.csv into)."""
        self.encodings_dir = os.path.join(log_dir, 'encodings')

    def path(self, iteration):
        return os.path.join(self.encodings_dir, f'{iteration}.enc')

    def create_interceptor(self, trainer):
        os.makedirs(self.encodings_dir, exist_ok=True)

        def intercept(it, _):
            encodings = np.asarray(trainer.network.encoding_prediction(trainer.data),
                                   dtype=np.float32)
            assert len(encodings) == len(trainer.data), \
                f'encodings + cells have different length: {len(encodings)} != {len(trainer.data)}'
            tmp = self.path(it) + '.tmp'
            with open(tmp, 'wb') as f:
                pickle.dump(encodings, f, protocol=pickle.HIGHEST_PROTOCOL)
            os.replace(tmp, self.path(it))       # readers list the directory while we train

        return intercept


def load_encodings(log_dir, iteration):
    """What import_encodings.load_coords / convert_encodings_to_mp4.load_coords read (before
    their x255)."""
    with open(EncodingFiles(log_dir).path(iteration), 'rb') as f:
        return pickle.load(f)


class Checkpoints:
    """Interceptor that saves `network.save_checkpoint()` every `every` iterations to
    `<log_dir>/checkpoint.npz` (atomic replace), and `resume(trainer)` that restores the newest
    one before a run.  The checkpoint also holds the iteration number and the host RNG states
    (numpy's global state = batch sampler, and the private prior generator), so
    `trainer.run(n, interceptor, start_iteration=ck.resume(trainer))` continues the
    interrupted run: same batches, same priors, iteration numbers that do not collide with the
    recorder's.  The reference persists nothing but encodings and loss CSVs; this
    is row f4 of its "next" list wired into the interceptor mechanism."""

    def __init__(self, log_dir, every=1):
        self.path = os.path.join(log_dir, 'checkpoint.npz')
        self.every = max(1, int(every))

    def resume(self, trainer):
        """-> the iteration to continue with (0 when there is no checkpoint; truthy otherwise)."""
        if not os.path.exists(self.path):
            return 0
        meta = trainer.network.load_checkpoint(self.path) or {}
        return int(meta.get('iteration', -1)) + 1

    def create_interceptor(self, trainer):
        os.makedirs(os.path.dirname(self.path) or '.', exist_ok=True)

        def intercept(it, _):
            if (it + 1) % self.every == 0:
                trainer.network.save_checkpoint(self.path, extra={'iteration': it})

        return intercept
