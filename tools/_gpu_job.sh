python tools/rms_tma_bench.py gpurun_out/r2_rms_l2_256.jsonl > gpurun_out/r2_rms_l2_256.log 2>&1; echo "rc=$?"
python - <<'P'
import json
for l in open('gpurun_out/r2_rms_l2_256.jsonl'):
    r=json.loads(l); print(r['K'],r['N'],r['batch'],'tma',r['tma_state'],'il',r['interleave'],'l2',r['l2_256'],round(r['ms'],3),round(r['GB/s']))
P
