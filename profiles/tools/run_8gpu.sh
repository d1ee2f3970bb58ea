set -x
cd /root/repo
nvidia-smi -L | wc -l
for N in 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r01_v10_bench_${N}gpu.json 2> gpurun_out/r01_v10_bench_${N}gpu.err
tail -2 gpurun_out/r01_v10_bench_${N}gpu.err
python profiles/tools/bench_summary.py gpurun_out/r01_v10_bench_${N}gpu.json
done
