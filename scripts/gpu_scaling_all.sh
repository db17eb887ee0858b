# Scaling artefacts on one 8-GPU box: gpurun --gpus 8 -- bash scripts/gpu_scaling_all.sh
set -x
mkdir -p gpurun_out
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_c4_n$n.err > gpurun_out/final_bench_n$n.json
done
python scripts/bench_c5_scaling.py --steps 5 --warmup 3 2>/dev/null > gpurun_out/final_c5_scaling.jsonl
for n in 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) scripts/bench_c5_scaling.py --steps 5 --warmup 3 2> gpurun_out/scale_c5_n$n.err >> gpurun_out/final_c5_scaling.jsonl
done
cut -c1-170 gpurun_out/final_bench_n*.json gpurun_out/final_c5_scaling.jsonl
