#!/bin/bash
# round 2, visit O: smoke() with the once-per-image stages; bench.py under torchrun on 2 GPUs (both arms), as the driver launches it
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke_r2o.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_r2o.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2_r2o.json 2> gpurun_out/bench_n2_r2o.err; echo "bench n2 exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2_r2o.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['dropin']['value'], d['once_per_image'], {k:v['value'] for k,v in d['configs'].items()})"
tail -n 3 gpurun_out/bench_n2_r2o.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2_r2o.json 2> gpurun_out/bench_ref_n2_r2o.err; echo "reference arm n2 exit $?"
cut -c1-400 gpurun_out/bench_ref_n2_r2o.json
