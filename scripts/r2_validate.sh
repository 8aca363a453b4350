# round-end validation: GPU test suite, smoke, the default bench line, the config-2 line
timeout -k 5 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2v_tests.log
timeout -k 5 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2v_smoke.log
timeout -k 5 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2v_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'], d['roofline']['frac'], d.get('multipathfinder_wall'))"
timeout -k 5 300 python bench.py --config cfg2_funnel100_p8_k1000_j6 --steps 20 --warmup 3 --no-wall --no-cpu-baseline > gpurun_out/r2v_bench_cfg2.json 2> gpurun_out/r2v_bench_cfg2.err; echo "cfg2 rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2v_bench_cfg2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
