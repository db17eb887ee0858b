set -x
nvidia-smi -L | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 2> gpurun_out/bench_n8_err.log > gpurun_out/bench_n8.json
cat gpurun_out/bench_n8.json | cut -c1-250
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_n8_err.log | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 3 --warmup 3 2> gpurun_out/bench_n4_err.log > gpurun_out/bench_n4.json
cat gpurun_out/bench_n4.json | cut -c1-250
