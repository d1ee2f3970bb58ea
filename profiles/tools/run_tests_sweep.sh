#!/bin/bash
# usage: run_tests_sweep.sh "ENV=.. ENV=.." ...   (GPU tests first, then one short bench per configuration)
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
bash profiles/tools/sweep.sh "$@"
