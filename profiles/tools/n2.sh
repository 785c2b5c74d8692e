# N=2 A/B of the collective: python bench.py under torchrun, one port per run
mkdir -p gpurun_out/n2
port=29600
for c in none peer nccl; do
  port=$((port+7))
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 200 --warmup 10 --collective $c 2>gpurun_out/n2/$c.err > gpurun_out/n2/$c.json
  echo "$c rc=$?"; cut -c1-200 gpurun_out/n2/$c.json; tail -4 gpurun_out/n2/$c.err | cut -c1-300
done
