# N=8 weak scaling with and without the NUMA binding of the ranks (bench.py --no-numa), plus the host topology
nproc > gpurun_out/r2_topo.txt; nvidia-smi topo -m >> gpurun_out/r2_topo.txt 2>&1; cat /sys/devices/system/node/node*/cpulist >> gpurun_out/r2_topo.txt
run() { # name, extra args, timeout, port
  timeout -k 5 $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 8 --steps 10 --warmup 3 --no-wall --no-mode-m --no-cpu-baseline $2 > gpurun_out/r2_$1.json 2> gpurun_out/r2_$1.err; echo "$1 rc=$?"; python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/r2_$1.json')); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('numa_bind_rank0'))
except Exception as e: print('$1 no json', e)"
}
run weak_n8_numa "" 240 29611
run weak_n8_nonuma "--no-numa" 240 29612
