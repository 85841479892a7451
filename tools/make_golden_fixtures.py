"""Regenerate tests/golden/ from the read-only reference checkout (run in the build container,
where /root/reference exists; the GPU box only sees the committed outputs).

  * example_matrix.mtx / example_barcodes.tsv / example_genes.tsv: the reference's own test
    fixtures (test/example_*), DATA files copied verbatim so the loader / recorder goldens of
    test/cell_type_training_test.py and test/db_recorder_test.py can be replayed.
  * loader_golden.json: what the reference's load_matrix body (pandas) returns for the
    fixture and for a synthetic COO with duplicates / zeros / gaps, plus sampler goldens.
"""
import json
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.environ.get("CELLCOMM_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def main():
    from oracle import loader_oracle as LO
    os.makedirs(OUT, exist_ok=True)
    for f in ("example_matrix.mtx", "example_barcodes.tsv", "example_genes.tsv"):
        shutil.copyfile(os.path.join(REF, "test", f), os.path.join(OUT, f))
    fx = os.path.join(OUT, "example_matrix.mtx")
    df = LO.load_matrix_pandas(fx)
    rng = np.random.default_rng(20260101)
    n = 600
    genes = rng.choice(np.arange(2, 90, 2), n)
    barcodes = rng.choice(np.r_[np.arange(3, 40), 77], n)
    vals = rng.integers(0, 30, n)
    syn = os.path.join(OUT, "synthetic_dups.mtx")
    with open(syn, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n%\n1 1 1\n")
        for g, b, v in zip(genes, barcodes, vals):
            f.write(f"{g} {b} {v}\n")
    ds = LO.load_matrix_pandas(syn)
    golden = {
        "fixture": {"index": df.index.tolist(), "columns": df.columns.tolist(),
                    "values": df.values.tolist(),
                    "sample_seed0_rows": df.sample(3, random_state=0).index.tolist()},
        "synthetic_dups": {"index": ds.index.tolist(), "columns": ds.columns.tolist(),
                           "values": ds.values.tolist(),
                           "sample_seed3_rows": ds.sample(8, random_state=3).index.tolist()},
    }
    with open(os.path.join(OUT, "loader_golden.json"), "w") as f:
        json.dump(golden, f)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
