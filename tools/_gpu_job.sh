python -m pytest tests -m gpu -q --tb=short --maxfail=20 > gpurun_out/r2_tests12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests12.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke12.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_tests12.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench12.json 2> gpurun_out/r2_bench12.err; echo "bench rc=$?" >> gpurun_out/r2_tests12.log
tail -n 5 gpurun_out/r2_tests12.log; tail -n 1 gpurun_out/r2_smoke12.log
