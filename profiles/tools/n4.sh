# N=4 A/B of the collective + the multi-GPU peer-exchange test
mkdir -p gpurun_out/n4
port=29700
for c in peer nccl none; do
  port=$((port+7))
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 4 --steps 200 --warmup 10 --collective $c 2>gpurun_out/n4/$c.err | grep '^{' > gpurun_out/n4/$c.json
  python -c "import json; d=json.load(open('gpurun_out/n4/$c.json')); print('$c', round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1))"
done
timeout 300 python -m pytest tests/test_peer_exchange.py -m gpu -x -q 2>&1 | tail -2
