set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
bash profiles/tools/sweep.sh "A=1" "DINT_B200_STRIP_CODEWORDS=10" "DINT_B200_UNIT_INTS=896 DINT_B200_UNIT_BYTES=1536"
python bench.py --type single_packed_dint --no-cpu --no-e2e --steps 5 > gpurun_out/v10_single.json 2>gpurun_out/v10_single.err; python profiles/tools/bench_summary.py gpurun_out/v10_single.json
