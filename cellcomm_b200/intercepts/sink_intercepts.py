"""`losses.csv` / `accuracy.csv` interceptors (API of the reference's
src/intercepts/sink_intercepts.py:6-31): both append one row per iteration through a
`DataSink` that is drained at interpreter exit.

    losses.csv    iteration,total-loss,g-loss,e-loss,d-loss
    accuracy.csv  iteration,pos-pct,neg-pct      (fractions of a freshly sampled batch that
                                                  the discriminator classifies correctly)
"""
import atexit

from .data_sink import DataSink

LOSS_COLUMNS = ('iteration', 'total-loss', 'g-loss', 'e-loss', 'd-loss')
ACCURACY_COLUMNS = ('iteration', 'pos-pct', 'neg-pct')


class SinkIntercepts:
    def __init__(self, log_dir):
        self.sink = DataSink(log_dir=log_dir)
        atexit.register(self.sink.drain_data)

    def _open(self, graph_id, columns):
        self.sink.add_graph_header(graph_id, list(columns))
        return lambda row: self.sink.add_data(graph_id, row)

    def save_losses(self):
        write = self._open('losses', LOSS_COLUMNS)

        def store_record(it, all_losses):
            # device-resident LossScalars become host floats here (one read per iteration)
            g, e, d = map(float, all_losses)
            write([it, g + e + d, g, e, d])

        return store_record

    def save_accuracy(self, trainer):
        write = self._open('accuracy', ACCURACY_COLUMNS)

        def intercept(it, _losses):
            batch = trainer.sample_cell_data()
            true_pos, true_neg = trainer.network.evaluate_discriminator_accuracy(batch)
            write([it, true_pos / len(batch), true_neg / len(batch)])

        return intercept
