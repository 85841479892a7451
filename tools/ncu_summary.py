"""Key metrics of an `ncu --set full` report as JSON (for profiles/):
    python tools/ncu_summary.py report.ncu-rep out.json "note" """
import csv
import io
import json
import subprocess
import sys

rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "launch__registers_per_thread",
        "launch__cluster_size", "sm__cycles_elapsed.avg.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
res = []
for r in rows[2:]:
    d = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            d[w] = (r[i] + " " + units[i]).strip()
    res.append(d)
json.dump({"note": note, "launches": res}, open(out, "w"), indent=1)
print(json.dumps(res[-1], indent=1))
