set -x
cd /root/repo
export PATH=$PATH:/usr/local/cuda/bin
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -x -q > gpurun_out/sanitizer_memcheck_all.log 2>&1; tail -5 gpurun_out/sanitizer_memcheck_all.log
timeout 900 compute-sanitizer --tool initcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_initcheck.log 2>&1; tail -4 gpurun_out/sanitizer_initcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests -m gpu -x -q -k "random_streams or query_log or interpolative" > gpurun_out/sanitizer_racecheck_tests.log 2>&1; tail -5 gpurun_out/sanitizer_racecheck_tests.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_synccheck.log 2>&1; tail -3 gpurun_out/sanitizer_synccheck.log
