set -x
cd /root/repo
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; head -14 gpurun_out/topo.txt; nproc; lscpu | grep -i "numa\|socket" | head
N=4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r01_final_bench_${N}gpu.json 2> gpurun_out/r01_final_bench_${N}gpu.err
tail -2 gpurun_out/r01_final_bench_${N}gpu.err
python profiles/tools/bench_summary.py gpurun_out/r01_final_bench_${N}gpu.json
