set -x
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 3 --warmup 3 2> gpurun_out/bench_n${n}_err.log > gpurun_out/bench_n$n.json
cut -c1-200 gpurun_out/bench_n$n.json
done
python -m pytest tests/test_robustness_gpu.py -q -k multi_device 2>&1 | tail -2
