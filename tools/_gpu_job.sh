CELLCOMM_DP_LOG=gpurun_out/r2_dp_check.jsonl python -m pytest tests/test_data_parallel_gpu.py -m gpu -q --tb=short -s > gpurun_out/r2_dp_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_dp_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/r2_dp_tests.log
$TR --master-port 29512 bench.py --workload record --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_record_n2.json 2> gpurun_out/r2_bench_record_n2.err; echo "record n2 rc=$?" >> gpurun_out/r2_dp_tests.log
$TR --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --strong --batch 2048 --no-roofline > gpurun_out/r2_bench_strong_n2.json 2> gpurun_out/r2_bench_strong_n2.err; echo "strong n2 rc=$?" >> gpurun_out/r2_dp_tests.log
tail -n 6 gpurun_out/r2_dp_tests.log
