"""The product entry point `python3 src` (= `python -m cellcomm_b200`; reference
src/__main__.py:44-66,89-99) on the GPU: load the 10x source, train, print / CSV the losses,
record every iteration's encodings -- single process and, with >= 2 GPUs, under torchrun (every
rank trains, rank 0 alone owns the log directory and the interceptors)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCE = "GSE122930_TAC_4_weeks_repA+B"        # DATA_SOURCES = SOURCES[1], src/__main__.py:30-40


def _write_data_dir(tmp, N=96, G=700, seed=5):
    rng = np.random.default_rng(seed)
    dense = (rng.random((N, G)) < 0.08) * rng.geometric(0.45, (N, G))
    dense[np.arange(N), rng.integers(0, G, N)] += 1
    dense[rng.integers(0, N, G), np.arange(G)] += 1
    with open(os.path.join(tmp, f"{SOURCE}_matrix.mtx"), "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate integer general\n%\n{G} {N} {(dense > 0).sum()}\n")
        for b in range(N):
            for g in np.flatnonzero(dense[b]):
                f.write(f"{g + 1} {b + 1} {int(dense[b, g])}\n")
    with open(os.path.join(tmp, f"{SOURCE}_barcodes.tsv"), "w") as f:
        f.writelines(f"BC{b:05d}-1\n" for b in range(N))
    with open(os.path.join(tmp, f"{SOURCE}_genes.tsv"), "w") as f:
        f.writelines(f"ENS{g:06d}\tsym{g}\n" for g in range(G))


def _run(cmd, tmp):
    env = dict(os.environ, CELLCOMM_DATA_DIR=tmp, CELLCOMM_ITERATIONS="2", CELLCOMM_BATCH_SIZE="16",
               PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "src"), ROOT]))
    out = subprocess.run(cmd, cwd=tmp, env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    return out.stdout


def _check_logs(tmp, stdout):
    lines = [l for l in stdout.splitlines() if " it: " in l and "TOT:" in l]
    assert len(lines) == 2, stdout[-1500:]                 # print_losses, one line per iteration
    runs = os.listdir(os.path.join(tmp, "logs"))
    assert len(runs) == 1 and runs[0].endswith("_test_e3")  # logs/<MM-DD-HHMM>_<RUN_ID>_e<Z>
    csv = open(os.path.join(tmp, "logs", runs[0], "losses.csv")).read().splitlines()
    assert csv[0] == "iteration,total-loss,g-loss,e-loss,d-loss" and len(csv) == 3
    assert all(np.isfinite([float(v) for v in row.split(",")]).all() for row in csv[1:])
    assert "importing cells ... DONE" in stdout             # DbRecorder.setup() ran (once)
    assert stdout.count("importing cells ... DONE") == 1


def test_python3_src_single_gpu(tmp_path):
    tmp = str(tmp_path)
    _write_data_dir(tmp)
    stdout = _run([sys.executable, os.path.join(ROOT, "src")], tmp)
    _check_logs(tmp, stdout)


def test_torchrun_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    tmp = str(tmp_path)
    _write_data_dir(tmp)
    port = 29700 + (os.getpid() % 200)
    stdout = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                   "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "cellcomm_b200"], tmp)
    _check_logs(tmp, stdout)
