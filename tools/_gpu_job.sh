python -m pytest tests -m gpu -q --tb=short --maxfail=20 > gpurun_out/r2_tests22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests22.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench22.json 2> gpurun_out/r2_bench22.err; echo "bench rc=$?" >> gpurun_out/r2_tests22.log
tail -n 5 gpurun_out/r2_tests22.log
