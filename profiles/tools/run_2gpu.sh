set -x
cd /root/repo
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q -k "block128 or query_log or sharded" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r01_v10_bench_2gpu.json 2> gpurun_out/r01_v10_bench_2gpu.err
tail -3 gpurun_out/r01_v10_bench_2gpu.err
python profiles/tools/bench_summary.py gpurun_out/r01_v10_bench_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r01_v10_bench_2gpu_reference.json 2> gpurun_out/r01_v10_bench_2gpu_reference.err
cut -c1-300 gpurun_out/r01_v10_bench_2gpu_reference.json
