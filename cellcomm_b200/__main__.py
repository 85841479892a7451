"""`python3 src` / `python3 -m cellcomm_b200` entry point (reference src/__main__.py:30-99):
train ContinuousCellBiGan on one of the GSE122930 sources with the print / CSV / MongoDB
interceptors, or `convert <matrix.mtx> <cells.csv>`.  Same module constants, same log-dir
rule (`logs/<MM-DD-HHMM>_<RUN_ID>_e<Z>` must not exist), SIGINT exits 0.

Extra, optional environment knobs (the reference has none): CELLCOMM_DATA_DIR (where the
`*_matrix.mtx`, `*_barcodes.tsv`, `*_genes.tsv` live; default `<repo>/data`),
CELLCOMM_ITERATIONS, CELLCOMM_BATCH_SIZE.
"""
import os
import pathlib
import signal
import sys
from datetime import datetime

from .cell_type_training import CellTraining, load_matrix
from .intercepts import (DbRecorder, SinkIntercepts, combined_interceptors, offset_iterations,
                         print_losses, skip_iterations)

DATA_DIR = os.environ.get('CELLCOMM_DATA_DIR',
                          os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'data'))


def data_file(file):
    return os.path.join(DATA_DIR, file)


def log_file(log_file_):
    return os.path.join('logs', log_file_)


def build_source(source_id):
    return {kind: data_file(f'{source_id}_{kind}.{ext}')
            for kind, ext in (('matrix', 'mtx'), ('barcodes', 'tsv'), ('genes', 'tsv'))}


SOURCE_IDS = [
    'GSE122930_TAC_1_week_repA+B',
    'GSE122930_TAC_4_weeks_repA+B',
    'GSE122930_Sham_1_week',
    'GSE122930_Sham_4_weeks_repA+B'
]

SOURCES = [build_source(src) for src in SOURCE_IDS]

RUN_ID = 'test'
DATA_SOURCES = SOURCES[1]
LOG_ID_TEMPLATE = '{}_' + RUN_ID + '_e{}'


def run_training(batch_size=128):
    data_source = load_matrix(DATA_SOURCES['matrix'], verbose=True)
    encoding_size = 3

    trainer = CellTraining(data_source, batch_size=batch_size, encoding_size=encoding_size)
    interceptors = create_interceptors(encoding_size, trainer, DATA_SOURCES)
    trainer.run(int(os.environ.get('CELLCOMM_ITERATIONS', '1')), interceptors)


def create_interceptors(encoding_size, trainer, sources):
    now = datetime.now().strftime('%m-%d-%H%M')
    full_run_id = LOG_ID_TEMPLATE.format(now, encoding_size)
    log_dir = log_file(full_run_id)
    check_log_dir(log_dir)

    sink = SinkIntercepts(log_dir)
    db_rec = DbRecorder(RUN_ID, sources)
    db_rec.setup()
    return combined_interceptors([
        print_losses(full_run_id),
        sink.save_losses(),
        offset_iterations(0, skip_iterations(1, db_rec.create_interceptor(trainer)))
    ])


def check_log_dir(log_dir):
    log_path = pathlib.Path(log_dir)
    if log_path.exists():
        raise AssertionError(f'duplicate run-id, log-dir: {log_dir}')
    log_path.mkdir(parents=True)


def store_converted_cell_file(matrix_file, cell_file):
    print('converting matrix file:')
    df = load_matrix(matrix_file, verbose=True)
    print('storing cell file:', cell_file, '... ', end='', flush=True)
    df.to_csv(cell_file)
    print('done')


def signal_handler(_, __):
    print('\tstopped')
    sys.exit(0)


def main(argv):
    signal.signal(signal.SIGINT, signal_handler)
    if len(argv) > 1:
        cmd = argv[1]
        if cmd == 'convert':
            assert len(argv) == 4, \
                'required parameters missing: convert <source-matrix-file> <convert-target-file>'
            store_converted_cell_file(argv[2], argv[3])
        else:
            print('unrecognised command:', cmd)
    else:
        run_training(int(os.environ.get('CELLCOMM_BATCH_SIZE', '128')))


if __name__ == '__main__':
    main(sys.argv)
