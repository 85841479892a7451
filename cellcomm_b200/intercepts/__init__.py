"""Drop-in for the reference's `src/intercepts` package: callback combinators and the loss
printer (reference __init__.py:8-38), plus the CSV sink, the MongoDB recorder and the plot
interceptors.  An interceptor is `Callable[[int, tuple[g, e, d]], None]`, fired once per
iteration by `CellTraining.run` (src/cell_type_training.py:40-50)."""
from datetime import datetime

from .plot_intercepts import PlotIntercepts
from .sink_intercepts import SinkIntercepts
from .db_recorder import DbRecorder
from .encoding_files import EncodingFiles


def combined_interceptors(interceptors):
    def call_all(it, losses):
        for ic in interceptors:
            ic(it, losses)

    return call_all


def skip_iterations(steps, interceptor):
    def intercept(it, losses):
        if (it % steps) >= (steps - 1):
            interceptor(it, losses)

    return intercept


def offset_iterations(offset, interceptor):
    def intercept(it, losses):
        if it >= offset:
            interceptor(it, losses)

    return intercept


def print_losses(full_run_id):
    def intercept(it, all_losses):
        # the losses may be device-resident LossScalars: one host read per iteration, here
        g_loss, e_loss, d_loss = (float(v) for v in all_losses)
        ts = datetime.now().strftime('%Y-%m-%d %H:%M:%S')
        print(f'[{ts}] {full_run_id} it: {it:6}  TOT: {g_loss + e_loss + d_loss:6.3f}  '
              f'G-L: {g_loss:6.3f}  E-L: {e_loss:6.3f}  D-L: {d_loss:6.3f}')

    return intercept
