"""CSV sink with batched appends (reference src/intercepts/data_sink.py:15-60): one
`<log_dir>/<graph>.csv` per graph id, a header line on registration, the same three
AssertionError messages."""
from typing import Any, Iterable

DEFAULT_LOG_DIR = 'logs'
DEFAULT_BATCH_SIZE = 1


def csv_line(values):
    return ','.join(str(s) for s in values) + '\n'


class DataSink:
    def __init__(self, log_dir=DEFAULT_LOG_DIR, batch_size=DEFAULT_BATCH_SIZE):
        self._log_dir = log_dir
        self._batch_size = batch_size
        self._graphs = {}      # graph id -> {'file', 'lines', 'size'}

    def add_graph_header(self, graph_id, fields: Iterable[Any]):
        if graph_id in self._graphs:
            raise AssertionError(f'duplicate graph name: {graph_id}')
        fields = list(fields)
        self._graphs[graph_id] = {'file': f'{self._log_dir}/{graph_id}.csv', 'lines': [],
                                  'size': len(fields)}
        self._append(graph_id, csv_line(fields))

    def add_data(self, graph_id, values: Iterable[Any]):
        if graph_id not in self._graphs:
            raise AssertionError(f'unknown graph: {graph_id}')
        graph = self._graphs[graph_id]
        if not len(values) == graph['size']:
            raise AssertionError(f'expected {graph["size"]} values, received: {values}')
        graph['lines'].append(csv_line(values))
        if len(graph['lines']) >= self._batch_size:
            self._drain(graph_id)

    def drain_data(self):
        for graph_id in self._graphs:
            self._drain(graph_id)

    def _drain(self, graph_id):
        lines = self._graphs[graph_id]['lines']
        if lines:
            self._append(graph_id, ''.join(lines))
            lines.clear()

    def _append(self, graph_id, text):
        with open(self._graphs[graph_id]['file'], 'a+') as f:
            f.write(text)
