# single-GPU bench lines of the other BASELINE configs with the final code + a launch list of config 5
# (shows that no library GEMM runs on the path: K8g is the only GEMM kernel)
for c in cfg2_funnel100_p8_k1000_j6 cfg4_hlogistic256_p32_k2000_j6 cfg5_dense4096_p16_k500_j10; do
  s=${c%%_*}
  timeout -k 5 400 python bench.py --config $c --steps 5 --warmup 3 --no-wall --no-cpu-baseline > gpurun_out/r2_bench_$s.json 2> gpurun_out/r2_bench_$s.err
  echo "$s rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/r2_bench_$s.json')); print('$s', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline'].get('frac'))"
done
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg5.csv python bench.py --config cfg5_dense4096_p16_k500_j10 --steps 1 --warmup 3 --no-wall --no-cpu-baseline --no-mode-m > gpurun_out/r2_launches_cfg5.log 2>&1
echo "ncu rc=$?"
