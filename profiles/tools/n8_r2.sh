# N=8 (round 2): configs[3] / configs[4] tests on 2, 4 and 8 ranks, then bench.py on 8 ranks: default (configs[1] weak
# scaling), shards (configs[3] as written) and mosaic (configs[4]).   gpurun --gpus 8 -- 'bash profiles/tools/n8_r2.sh'
mkdir -p gpurun_out/n8r2
timeout 420 python -m pytest tests/test_multi_gpu_configs.py tests/test_peer_exchange.py tests/test_gpu_parity.py -k "multi or peer or configs or nodata_tiles" -m gpu -x -q > gpurun_out/n8r2/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/n8r2/pytest.log
port=29700
for c in "granule" "shards" "mosaic" "tiles"; do
  port=$((port+7))
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 50 --warmup 5 --config $c 2>gpurun_out/n8r2/$c.err | grep '^{' > gpurun_out/n8r2/$c.json
  python -c "import json; d=json.load(open('gpurun_out/n8r2/$c.json')); print('$c:', round(d['ms_per_step'],4), 'ms', round(d['value'],1), 'Mpix/s', (round(d['e2e']['value'],1), round(d['e2e']['frac_of_h2d_peak'],3), round(d['e2e']['h2d_peak_gbs'],1)) if 'e2e' in d else '', d.get('peak_hbm_gb'))" || tail -5 gpurun_out/n8r2/$c.err
done
