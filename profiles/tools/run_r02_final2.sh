set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final2_pytest.log; tail -3 gpurun_out/r02_final2_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_final2_bench.json 2> gpurun_out/r02_final2_bench.err
timeout 600 python bench.py --format index > gpurun_out/r02_final2_bench_index.json 2> gpurun_out/r02_final2_bench_index.err
timeout 600 python bench.py --type single_packed_dint --no-cpu --no-e2e-out > gpurun_out/r02_final2_bench_single_packed.json 2> gpurun_out/r02_final2_bench_single_packed.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final2_launches.csv python bench.py --steps 2 --warmup 3 --replicas 32 --no-cpu --no-e2e > gpurun_out/ncu_l2.log 2>&1
python -c "
import json
for f in ('r02_final2_bench','r02_final2_bench_index','r02_final2_bench_single_packed'):
    try:
        r=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, r['value'], (r.get('prepass') or {}).get('ms'), (r.get('e2e_cold') or {}).get('value'), (r.get('e2e') or {}).get('value'))
    except Exception as e: print(f,'ERR',e)
"
