run() { # name, nproc, extra args, timeout, port
  if [ "$2" = "1" ]; then
    timeout -k 5 $4 python bench.py --gpus 1 --steps 10 --warmup 3 --no-wall --no-mode-m --no-cpu-baseline $3 > gpurun_out/r2_$1.json 2> gpurun_out/r2_$1.err
  else
    timeout -k 5 $4 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $5 bench.py --gpus $2 --steps 10 --warmup 3 --no-wall --no-mode-m --no-cpu-baseline $3 > gpurun_out/r2_$1.json 2> gpurun_out/r2_$1.err
  fi
  echo "$1 rc=$?"; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2_$1.json')); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms']['total'])
except Exception as e: print('$1 no json', e)"
}
run weak_n1 1 "" 200 0
run weak_n2 2 "" 200 29611
run strong_n2 2 "--scaling strong" 200 29612
run weak_n4 4 "" 200 29613
run strong_n4 4 "--scaling strong" 200 29614
