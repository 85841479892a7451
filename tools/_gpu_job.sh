python -m pytest tests/test_gemm_gpu.py tests/test_parity_baseline_shape_gpu.py tests/test_parity_gpu.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/r2_tests15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests15.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench15.json 2> gpurun_out/r2_bench15.err; echo "bench rc=$?" >> gpurun_out/r2_tests15.log
tail -n 5 gpurun_out/r2_tests15.log
