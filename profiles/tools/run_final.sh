set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01_final_pytest.log; tail -3 gpurun_out/r01_final_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r01_final_bench.json 2> gpurun_out/r01_final_bench.err
timeout 300 python bench.py --impl reference > gpurun_out/r01_final_bench_reference.json 2> gpurun_out/r01_final_bench_reference.err
python profiles/tools/bench_summary.py gpurun_out/r01_final_bench.json
python -c "
import json
d=json.loads(open('gpurun_out/r01_final_bench.json').read().strip().splitlines()[-1]); print(d['clocks'], d['gpu_launches'], d['roofline']['traffic'], d['prepass']['ms'])
r=json.loads(open('gpurun_out/r01_final_bench_reference.json').read().strip().splitlines()[-1]); print(r['value'], r['ms_per_step'])"
