python -m pytest tests -m gpu -q --tb=short --maxfail=20 > gpurun_out/r2_tests13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests13.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; echo "bench rc=$?" >> gpurun_out/r2_tests13.log
CELLCOMM_B200_NARROW_ELEMS=0 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench13_nonarrow.json 2> gpurun_out/r2_bench13_nonarrow.err
CELLCOMM_B200_NARROW_ELEMS=1048576 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench13_1m.json 2> gpurun_out/r2_bench13_1m.err
tail -n 5 gpurun_out/r2_tests13.log
