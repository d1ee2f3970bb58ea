set -x
cd /root/repo
export PATH=$PATH:/usr/local/cuda/bin
timeout 900 compute-sanitizer --tool memcheck --kernel-regex kns=dintk --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_memcheck.log 2>&1; tail -5 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python __graft_entry__.py smoke > gpurun_out/sanitizer_racecheck.log 2>&1; tail -8 gpurun_out/sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests -m gpu -x -q -k "edge_case or unaligned or fused_docids" > gpurun_out/sanitizer_memcheck_tests.log 2>&1; tail -6 gpurun_out/sanitizer_memcheck_tests.log
