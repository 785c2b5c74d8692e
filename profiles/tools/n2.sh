# N=2: collective A/B + the peer-exchange test; python bench.py under torchrun, one port per run
mkdir -p gpurun_out/n2
port=29600
for c in "peer" "peer --no-graph" "nccl" "none --no-graph"; do
  port=$((port+7)); tag=$(echo $c | tr -d ' -')
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 200 --warmup 10 --collective $c 2>gpurun_out/n2/$tag.err | grep '^{' > gpurun_out/n2/$tag.json
  python -c "import json; d=json.load(open('gpurun_out/n2/$tag.json')); print('$c:', round(d['ms_per_step'],4), round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['kernel_ms'],4))" || tail -5 gpurun_out/n2/$tag.err
done
timeout 300 python -m pytest tests/test_peer_exchange.py -m gpu -x -q 2>&1 | tail -2
