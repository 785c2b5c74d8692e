# N=8: the driver's scaling run (default flags) + the peer-exchange test on 8 GPUs
mkdir -p gpurun_out/n8
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/n8/nvsmi.txt 2>&1
port=29800
for c in peer nccl; do
  port=$((port+7))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 100 --warmup 5 --collective $c 2>gpurun_out/n8/$c.err | grep '^{' > gpurun_out/n8/$c.json
  python -c "import json; d=json.load(open('gpurun_out/n8/$c.json')); print('$c', round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1))" || tail -5 gpurun_out/n8/$c.err
done
timeout 200 python -m pytest tests/test_peer_exchange.py -m gpu -x -q 2>&1 | tail -2
