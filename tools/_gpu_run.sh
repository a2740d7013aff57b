mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_r1_final.err
python -c "
import json
r = json.load(open('gpurun_out/bench_r1_final.json'))
print('ms/step', r['ms_per_step'], 'e2e ms', r['e2e']['ms_per_step'], 'frac', r['roofline']['frac'], 'launches', r['gpu_launches'])
"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain_short.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file gpurun_out/launches_sell.csv $CMD > gpurun_out/ncu_list_sell.log 2>&1
echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list_sell.log; wc -l gpurun_out/launches_sell.csv
ncu --set full --clock-control none --import-source on -k regex:k_sell_stream -s 400 -c 12 -o gpurun_out/prof_bench_sell $CMD > gpurun_out/ncu_full_sell.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_sell.log
gzip -f gpurun_out/launches_sell.csv
