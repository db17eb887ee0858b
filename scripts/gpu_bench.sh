set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2> gpurun_out/bench_n2_err.log | tee gpurun_out/bench_n2.json | cut -c1-250
tail -5 gpurun_out/bench_n2_err.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>> gpurun_out/bench_n2_err.log | cut -c1-250
