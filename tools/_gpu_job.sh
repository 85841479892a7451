set -x
python tools/wgrad_fused_one.py 6738 33694 128,2048 2 --blocked > gpurun_out/r2_blocked_one.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:persistent -o gpurun_out/r2_blocked_full -f python tools/wgrad_fused_one.py 6738 33694 128,2048 2 --blocked > gpurun_out/r2_blocked_ncu.log 2>&1
tail -3 gpurun_out/r2_blocked_one.log; tail -3 gpurun_out/r2_blocked_ncu.log
