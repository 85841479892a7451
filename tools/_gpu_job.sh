python -m pytest tests -m gpu -q --tb=short --maxfail=30 --durations=8 > gpurun_out/r2_tests7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests7.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke7.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_tests7.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc=$?" >> gpurun_out/r2_tests7.log
tail -n 6 gpurun_out/r2_tests7.log
