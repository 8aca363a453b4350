# config 5 as BASELINE defines it (128 paths over 8 GPUs) with the final code (K8g, no library GEMM)
timeout -k 5 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 --no-wall --no-mode-m --no-cpu-baseline --config cfg5_dense4096_p16_k500_j10 > gpurun_out/r2_cfg5_n8_k8g.json 2> gpurun_out/r2_cfg5_n8_k8g.err; echo "rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2_cfg5_n8_k8g.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stage_ms'])"
