"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum` pass over tools/one_step.py (one trainings_step):

    python tools/summarize_traffic.py launches.csv profiles/r01_gemm_traffic.json \
        profiles/r01_ncu_launch_summary.json

Writes (1) DRAM bytes per launch of the two GEMM sets bench.py reports rooflines for (the
tensor-bound forward / dgrad GEMMs and the HBM-bound wgrad GEMMs with the fused RMSprop
epilogue), (2) per-kernel time / traffic shares of the step.
"""
import csv
import io
import json
import sys
from collections import defaultdict


def unit_scale(unit):
    u = unit.strip().lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1.0, "us": 1e3,
            "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "second": 1e9,
            "s": 1e9}.get(u, 1.0)


def main():
    src, out_traffic, out_summary = sys.argv[1:4]
    lines = open(src).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    per_launch = defaultdict(dict)
    for r in rows:
        val = float(r["Metric Value"].replace(",", "")) * unit_scale(r["Metric Unit"])
        per_launch[int(r["ID"])]["name"] = r["Kernel Name"]
        per_launch[int(r["ID"])][r["Metric Name"]] = val
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    sets = {"tensor": [0, 0.0, 0.0], "fused": [0, 0.0, 0.0]}
    for _, m in sorted(per_launch.items()):
        name = m["name"]
        short = name.split("(")[0].replace("void ", "").replace("cc::", "")
        t = m.get("gpu__time_duration.sum", 0.0)
        by = m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a = agg[short]
        a[0] += 1
        a[1] += t
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
        if "gemm_tcgen05" in short:
            # the fused-optimiser wgrad is the only instantiation with 8 epilogue warps
            key = "fused" if ", 8, " in short else "tensor"
            sets[key][0] += 1
            sets[key][1] += t
            sets[key][2] += by
    total_t = sum(a[1] for a in agg.values())
    summary = [{"kernel": k, "launches": a[0], "ms": a[1] / 1e6, "share_pct": 100 * a[1] / total_t,
                "dram_read_GB": a[2] / 1e9, "dram_write_GB": a[3] / 1e9}
               for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    with open(out_summary, "w") as f:
        json.dump({"step_ms_serialised": total_t / 1e6, "kernels": summary}, f, indent=1)
    traffic = {
        "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the "
                "launches of one trainings_step from one ncu pass "
                "(--clock-control none); source: " + src.split("/")[-1],
        "tensor_launches": sets["tensor"][0],
        "tensor_bytes_per_launch": sets["tensor"][2] / max(sets["tensor"][0], 1),
        "tensor_ms_serialised": sets["tensor"][1] / 1e6,
        "fused_launches": sets["fused"][0],
        "fused_bytes_per_launch": sets["fused"][2] / max(sets["fused"][0], 1),
        "fused_ms_serialised": sets["fused"][1] / 1e6,
        "gemm_share_of_step_serialised": (sets["tensor"][1] + sets["fused"][1]) / total_t,
    }
    with open(out_traffic, "w") as f:
        json.dump(traffic, f, indent=1)
    print(json.dumps(traffic, indent=1))
    for s in summary[:14]:
        print(f"{s['ms']:8.3f} ms {s['share_pct']:5.1f}%  n={s['launches']:4d}  "
              f"R {s['dram_read_GB']:6.2f} GB W {s['dram_write_GB']:6.2f} GB  {s['kernel'][:80]}")


if __name__ == "__main__":
    main()
