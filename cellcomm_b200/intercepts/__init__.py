"""Drop-in for the reference's `src/intercepts` package.

An interceptor is `Callable[[int, tuple[g, e, d]], None]`; `CellTraining.run` fires it once per
iteration with the iteration number and the three summed losses
(src/cell_type_training.py:40-50).  This module provides the four helpers callers import from
the package (reference src/intercepts/__init__.py:8-38) -- fan-out, two iteration filters and
the stdout loss line -- plus the sink / recorder / plot / encoding-file interceptor factories.

The filters share one gate: an interceptor wrapped with a predicate on the iteration number.
"""
import time

from .plot_intercepts import PlotIntercepts
from .sink_intercepts import SinkIntercepts
from .db_recorder import DbRecorder
from .encoding_files import Checkpoints, EncodingFiles

__all__ = ["combined_interceptors", "skip_iterations", "offset_iterations", "print_losses",
           "PlotIntercepts", "SinkIntercepts", "DbRecorder", "EncodingFiles", "Checkpoints"]

LOSS_LINE = ("[{stamp}] {run} it: {it:6}  TOT: {total:6.3f}  G-L: {g:6.3f}  E-L: {e:6.3f}  "
             "D-L: {d:6.3f}")


class _Gated:
    """Forward (iteration, losses) to `target` when `wanted(iteration)` holds."""

    __slots__ = ("wanted", "target")

    def __init__(self, wanted, target):
        self.wanted, self.target = wanted, target

    def __call__(self, iteration, losses):
        if self.wanted(iteration):
            self.target(iteration, losses)


def combined_interceptors(interceptors):
    """One interceptor that calls every given one, in order (reference :8-13)."""
    chain = tuple(interceptors)

    def fan_out(iteration, losses):
        for receiver in chain:
            receiver(iteration, losses)

    return fan_out


def skip_iterations(steps, interceptor):
    """Fire on the LAST iteration of every block of `steps`: steps-1, 2*steps-1, ...
    (reference :16-21)."""
    return _Gated(lambda iteration: iteration % steps == steps - 1, interceptor)


def offset_iterations(offset, interceptor):
    """Fire from iteration `offset` on (reference :24-29)."""
    return _Gated(lambda iteration: iteration >= offset, interceptor)


def print_losses(full_run_id):
    """Timestamped loss line on stdout, same text as the reference's (:32-38).  The losses may
    be device-resident `LossScalar`s: this is where they are read back, once per iteration."""

    def report(iteration, all_losses):
        g, e, d = (float(v) for v in all_losses)
        print(LOSS_LINE.format(stamp=time.strftime('%Y-%m-%d %H:%M:%S'), run=full_run_id,
                               it=iteration, total=g + e + d, g=g, e=e, d=d))

    return report
