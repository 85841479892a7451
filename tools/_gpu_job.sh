python -m pytest tests/test_gemm_gpu.py tests/test_parity_gpu.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/r2_tests18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests18.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench18.json 2> gpurun_out/r2_bench18.err; echo "bench rc=$?" >> gpurun_out/r2_tests18.log
tail -n 4 gpurun_out/r2_tests18.log
