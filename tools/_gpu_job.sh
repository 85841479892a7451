python -m pytest tests -m gpu -q --tb=short --maxfail=20 > gpurun_out/r2_tests9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests9.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?" >> gpurun_out/r2_tests9.log
tail -n 4 gpurun_out/r2_tests9.log
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench9.json'))
print('ms', d['ms_per_step'], 'launches/step', d['gpu_launches']/10, 'tensor', d['roofline']['ms_per_step'], d['roofline']['frac'], 'hbm', d['roofline_hbm']['ms_per_step'], 'b128', d['reference_batch']['ms_per_step'], d['reference_batch']['e2e']['ms_per_step'])
P
