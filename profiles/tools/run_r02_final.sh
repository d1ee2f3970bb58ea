set -x
cd /root/repo
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest.log; tail -3 gpurun_out/r02_final_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r02_final_bench.json 2> gpurun_out/r02_final_bench.err
timeout 600 python bench.py --format index > gpurun_out/r02_final_bench_index.json 2> gpurun_out/r02_final_bench_index.err
timeout 300 python bench.py --impl reference > gpurun_out/r02_final_bench_reference.json 2> gpurun_out/r02_final_bench_reference.err
timeout 300 python bench.py --impl reference --format index --steps 2 > gpurun_out/r02_final_bench_reference_index.json 2> gpurun_out/r02_final_bench_reference_index.err
timeout 300 python profiles/tools/small_configs.py > gpurun_out/r02_final_small_configs.json 2> gpurun_out/r02_final_small_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 2 --warmup 3 --replicas 32 --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_index_launches.csv python bench.py --format index --steps 2 --warmup 3 --replicas 32 --no-cpu > gpurun_out/ncu_li.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_strips -s 6 -c 2 -o gpurun_out/r02_final_strips -f python bench.py --steps 1 --warmup 3 --replicas 32 --no-cpu --no-e2e > gpurun_out/ncu_f.log 2>&1
tail -2 gpurun_out/ncu_f.log
