python -m pytest tests -m gpu -q --tb=short --maxfail=20 > gpurun_out/r2_tests14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests14.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench14.json 2> gpurun_out/r2_bench14.err; echo "bench rc=$?" >> gpurun_out/r2_tests14.log
CELLCOMM_B200_BLOCKED_STATE=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench14_rows.json 2> gpurun_out/r2_bench14_rows.err
tail -n 5 gpurun_out/r2_tests14.log
