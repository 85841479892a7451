python -m pytest tests/test_gemm_gpu.py -m gpu -q --tb=line > gpurun_out/r2_tests6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests6.log
CC_GEMM_PAIR=1 python tools/gemm_bench.py 2048 gpurun_out/r2_gemm_bench_pair1b.json > gpurun_out/r2_gemm_bench_pair1b.log 2>&1
CC_GEMM_PAIR=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench6_pair1.json 2> gpurun_out/r2_bench6_pair1.err
python -m pytest tests/test_api_gpu.py tests/test_parity_gpu.py -m gpu -q --tb=short > gpurun_out/r2_tests6b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests6b.log
tail -n 3 gpurun_out/r2_tests6.log gpurun_out/r2_tests6b.log
