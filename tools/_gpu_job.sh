rm -f gpurun_out/parity_r2b.jsonl
CELLCOMM_PARITY_LOG=gpurun_out/parity_r2b.jsonl python -m pytest tests/test_parity_gpu.py tests/test_parity_baseline_shape_gpu.py -m gpu -q --tb=short > gpurun_out/r2_tests8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests8.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err; echo "bench rc=$?" >> gpurun_out/r2_tests8.log
CMD="python bench.py --steps 2 --warmup 3 --graph 0 --no-cpu-baseline --small-batch 0 --dense-e2e-steps 0 --no-roofline"
$CMD > gpurun_out/r2_plain_launches.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:cc:: -s 1700 -c 2600 --csv --log-file gpurun_out/r2_ncu_launches_b2048.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2_tests8.log
tail -n 5 gpurun_out/r2_tests8.log
