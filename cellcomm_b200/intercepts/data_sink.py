"""CSV sink behind `SinkIntercepts` (API of the reference's src/intercepts/data_sink.py:15-60).

Every graph id owns one file `<log_dir>/<graph id>.csv`: registering the graph writes the
header row, `add_data` buffers a row and flushes once `batch_size` rows are waiting,
`drain_data` flushes everything (the owner hooks it to `atexit`).  The three AssertionError
texts callers can see are the reference's.
"""
from typing import Any, Iterable

DEFAULT_LOG_DIR = 'logs'
DEFAULT_BATCH_SIZE = 1


def csv_line(values):
    """One CSV row: `str()` of every value, comma separated, newline terminated."""
    return ','.join(map(str, values)) + '\n'


class _Series:
    """One CSV file: fixed column count, rows buffered until flushed."""

    def __init__(self, path, columns):
        self.path, self.columns, self.pending = path, columns, []

    def write(self, text):
        with open(self.path, 'a+') as out:
            out.write(text)

    def flush(self):
        if self.pending:
            self.write(''.join(self.pending))
            self.pending = []


class DataSink:
    def __init__(self, log_dir=DEFAULT_LOG_DIR, batch_size=DEFAULT_BATCH_SIZE):
        self._log_dir, self._batch_size = log_dir, batch_size
        self._series = {}

    def add_graph_header(self, graph_id, fields: Iterable[Any]):
        if graph_id in self._series:
            raise AssertionError(f'duplicate graph name: {graph_id}')
        names = list(fields)
        series = _Series(f'{self._log_dir}/{graph_id}.csv', len(names))
        self._series[graph_id] = series
        series.write(csv_line(names))

    def add_data(self, graph_id, values: Iterable[Any]):
        series = self._series.get(graph_id)
        if series is None:
            raise AssertionError(f'unknown graph: {graph_id}')
        if len(values) != series.columns:
            raise AssertionError(f'expected {series.columns} values, received: {values}')
        series.pending.append(csv_line(values))
        if len(series.pending) >= self._batch_size:
            series.flush()

    def drain_data(self):
        for series in self._series.values():
            series.flush()
