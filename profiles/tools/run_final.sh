set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r01_final_pytest.log; tail -3 gpurun_out/r01_final_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r01_final_bench.json 2> gpurun_out/r01_final_bench.err
timeout 300 python bench.py --type single_packed_dint --no-cpu --no-e2e > gpurun_out/r01_final_bench_single.json 2> gpurun_out/r01_final_bench_single.err
timeout 300 python bench.py --impl reference > gpurun_out/r01_final_bench_reference.json 2> gpurun_out/r01_final_bench_reference.err
python profiles/tools/bench_summary.py gpurun_out/r01_final_bench.json gpurun_out/r01_final_bench_single.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_final_launches.csv python bench.py --steps 2 --warmup 3 --replicas 32 --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_strips -s 2 -c 2 -o gpurun_out/r01_final_strips python bench.py --steps 1 --warmup 3 --replicas 32 --no-cpu --no-e2e > gpurun_out/ncu_f.log 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/r01_final_bench.json').read().strip().splitlines()[-1]); print(d['clocks'], d['gpu_launches'], d['e2e']['value'], d['cpu_baseline']['value'])
r=json.loads(open('gpurun_out/r01_final_bench_reference.json').read().strip().splitlines()[-1]); print(r['value'], r['ms_per_step'])"
