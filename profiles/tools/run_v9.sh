set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
bash profiles/tools/sweep.sh "A=1" "DINT_B200_UNIT_INTS=768 DINT_B200_UNIT_BYTES=1280" "DINT_B200_STRIP_WARPS=32 DINT_B200_UNIT_INTS=768 DINT_B200_UNIT_BYTES=1280"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_strips -s 2 -c 2 -o gpurun_out/r01_v9_strips python bench.py --steps 1 --warmup 1 --replicas 32 --no-cpu --no-e2e > gpurun_out/ncu_f.log 2>&1
