N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
L=gpurun_out/r2_n${N}.log; : > $L
if [ "$N" = "8" ]; then
CELLCOMM_DP_LOG=gpurun_out/r2_dp_check_n8.jsonl python -m pytest tests/test_data_parallel_gpu.py -m gpu -q --tb=short -k "hardware and 8" > gpurun_out/r2_dp_tests_n8.log 2>&1; echo "dp pytest rc=$?" >> $L
$TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_train_n$N.json 2> gpurun_out/r2_bench_train_n$N.err; echo "train rc=$?" >> $L
$TR --master-port 29522 bench.py --workload record --gpus $N --steps 2 --warmup 1 > gpurun_out/r2_bench_record_n$N.json 2> gpurun_out/r2_bench_record_n$N.err; echo "record rc=$?" >> $L
$TR --master-port 29523 bench.py --gpus $N --steps 10 --warmup 3 --strong --batch 2048 --no-roofline > gpurun_out/r2_bench_strong_n$N.json 2> gpurun_out/r2_bench_strong_n$N.err; echo "strong rc=$?" >> $L
fi
$TR --master-port 29524 bench.py --workload classify --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_classify_n$N.json 2> gpurun_out/r2_bench_classify_n$N.err; echo "classify rc=$?" >> $L
$TR --master-port 29525 bench.py --workload encode --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_encode_n$N.json 2> gpurun_out/r2_bench_encode_n$N.err; echo "encode rc=$?" >> $L
cat $L
