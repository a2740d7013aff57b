mkdir -p gpurun_out
python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -c 500 gpurun_out/bench_n2.err
python -c "
import json
r = json.loads([l for l in open('gpurun_out/bench_n2.json') if l.startswith('{')][-1])
print('n_gpus', r['n_gpus'], 'value', r['value'], 'ms/step', r['ms_per_step'], 'e2e', r['e2e']['value'], 'frac', r['roofline']['frac'])
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>&1 | tail -2 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dd_bench.py --refine 3 --replicate-below 200000 2>&1 | grep "^{" | tail -1 > gpurun_out/dd2_r3_sell.json; cut -c1-700 gpurun_out/dd2_r3_sell.json
