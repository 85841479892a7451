python bench.py --workload classify --steps 5 --warmup 3 > gpurun_out/r2_bench_classify_n1.json 2> gpurun_out/r2_bench_classify_n1.err; echo "classify rc=$?"
python bench.py --workload encode --steps 3 --warmup 3 > gpurun_out/r2_bench_encode_n1.json 2> gpurun_out/r2_bench_encode_n1.err; echo "encode rc=$?"
python bench.py --workload record --steps 2 --warmup 1 > gpurun_out/r2_bench_record_n1.json 2> gpurun_out/r2_bench_record_n1.err; echo "record rc=$?"
