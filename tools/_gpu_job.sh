CELLCOMM_BENCH_GEMM_TABLE=gpurun_out/r2_gemm_table_b128_floor.txt python bench.py --batch 128 --small-batch 0 --steps 10 --warmup 3 --no-cpu-baseline --dense-e2e-steps 0 > gpurun_out/r2_bench_b128_floor.json 2> gpurun_out/r2_bench_b128_floor.err; echo "rc=$?"
head -16 gpurun_out/r2_gemm_table_b128_floor.txt
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_b128_floor.json'))
print('ms', d['ms_per_step'], 'tensor set', d['roofline']['ms_per_step'], 'hbm set', d['roofline_hbm']['ms_per_step'], 'gemm total', d['roofline']['gemm_ms_per_step'])
P
