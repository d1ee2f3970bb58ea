set -x
cd /root/repo
export PATH=$PATH:/usr/local/cuda/bin
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_discover.py tests/test_gpu_queries.py -x -q > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -6 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python -m pytest tests/test_gpu_discover.py tests/test_gpu_queries.py -x -q -k "edge_streams or hints or index_scan_with or equal_reference" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -6 gpurun_out/r02_sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/r02_sanitizer_initcheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_initcheck.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python -m pytest tests/test_gpu_discover.py -x -q -k "hints or index_scan_with" > gpurun_out/r02_sanitizer_synccheck.log 2>&1; tail -4 gpurun_out/r02_sanitizer_synccheck.log
