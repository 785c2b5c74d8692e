# N=2: the peer exchange, and its collective fallback to NCCL when a rank cannot set its block up (fault injected)
mkdir -p gpurun_out/n2f
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/n2f/peer.err | grep '^{' > gpurun_out/n2f/peer.json
python -c "import json; d=json.load(open('gpurun_out/n2f/peer.json')); print('peer', round(d['ms_per_step'],4), d['config']['collective'][:40])"
HSR_PEER_FAIL_RANK=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29917 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/n2f/fallback.err | grep '^{' > gpurun_out/n2f/fallback.json
python -c "import json; d=json.load(open('gpurun_out/n2f/fallback.json')); print('fallback', round(d['ms_per_step'],4), d['config']['collective'][:40])"
grep "peer exchange unavailable" gpurun_out/n2f/fallback.err | head -2
timeout 200 python -m pytest tests/test_peer_exchange.py -m gpu -x -q 2>&1 | tail -2
