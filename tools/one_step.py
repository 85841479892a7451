"""Profiling target: ONE eager trainings_step of the bench workload between
cudaProfilerStart/Stop (use with `ncu --profile-from-start off`).

    python tools/one_step.py [batch] [cells]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from cellcomm_b200 import ops  # noqa: E402
from cellcomm_b200.cell_type_training import CellTraining  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
dev = torch.device("cuda", 0)
data = bench.make_matrix(cells, bench.GENES, 20260101, dev)
np.random.seed(0)
trainer = CellTraining(data, batch_size=B, encoding_size=bench.Z)
e = trainer.network._engine
rowptr, colidx, values = data.device_csr(dev)
x16 = ops.alloc2d(B, bench.GENES, device=dev)
e.reserve(B)


def step():
    idx = torch.from_numpy(np.random.permutation(cells)[:B]).to(dev)
    ops.gather_rows(rowptr, colidx, values, bench.GENES, row_idx=idx, out16=x16)
    e.draw_latents(B)
    out = e.train_step(x16)
    e.join()
    return out


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
losses = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("losses", [float(v) for v in losses])
