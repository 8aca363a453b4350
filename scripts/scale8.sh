run() { # name, extra args
  timeout -k 5 $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 8 --steps 10 --warmup 3 --no-wall --no-mode-m --no-cpu-baseline $2 > gpurun_out/r2_$1.json 2> gpurun_out/r2_$1.err; echo "$1 rc=$?"; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2_$1.json')); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])
except Exception as e: print('$1 no json', e)"
}
run weak_n8 "" 240 29601
run strong_n8 "--scaling strong" 240 29602
run cfg5_n8 "--config cfg5_dense4096_p16_k500_j10" 400 29603
