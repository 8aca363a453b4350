run() { # name, extra args, timeout, port
  timeout -k 5 $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 8 --steps 10 --warmup 3 $2 > gpurun_out/r2b_$1.json 2> gpurun_out/r2b_$1.err; echo "$1 rc=$?"; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2b_$1.json')); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms']['total'])
except Exception as e: print('$1 no json', e)"
}
run weak_n8 "" 240 29601
run strong_n8 "--scaling strong" 240 29602
