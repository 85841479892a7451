CC_GEMM_RMS_WARPS=12 python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=line -k "fused_rmsprop or wgrad_with_fused" > gpurun_out/r2_tests11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests11.log
tail -n 3 gpurun_out/r2_tests11.log
python tools/rms_tma_bench.py gpurun_out/r2_rms_w12.jsonl > gpurun_out/r2_rms_w12.log 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2_rms_w12.jsonl'):
    r=json.loads(l); print(r['K'],r['N'],r['batch'],'tma',r['tma_state'],'warps',r['epi_warps'],round(r['ms'],3),round(r['GB/s']))
P
