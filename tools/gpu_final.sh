#!/bin/bash
# final visit of a session: whole gpu suite, smoke, the default bench exactly as the driver runs it (timed), reference arm,
# ncu launch list + DRAM traffic of one cfg3 step and of one VAE decode / T5 encode with the final build
mkdir -p gpurun_out
R=${1:-r2t}
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=300 --timeout-method=thread > gpurun_out/pytest_gpu_$R.log 2>&1; echo "pytest exit $?"
grep -E "passed|failed|error" gpurun_out/pytest_gpu_$R.log | tail -n 2; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu_$R.log | head
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$R.log 2>&1; echo "smoke exit $?"; tail -n 2 gpurun_out/smoke_$R.log | cut -c1-250
S0=$(date +%s); timeout 1500 python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; echo "bench exit $? in $(( $(date +%s) - S0 )) s"
python -c "
import json; d=json.load(open('gpurun_out/bench_$R.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['dropin']['value'], {k:v['value'] for k,v in d['configs'].items()}, d['gpu_eager_baseline']['value'], d['cpu_baseline'], d['once_per_image'], d['clocks'])"
tail -n 2 gpurun_out/bench_$R.err
S0=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/bench_ref_$R.json 2> gpurun_out/bench_ref_$R.err; echo "ref exit $? in $(( $(date +%s) - S0 )) s"
cut -c1-300 gpurun_out/bench_ref_$R.json
KREG='regex:tcgen05|ln_modulate|gemv_kernel|rope_table|timestep_embed|set_float|mod_cache'
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREG" --csv --log-file gpurun_out/launches_${R}_cfg3.csv \
   python tools/one_step.py --workload cfg3 > gpurun_out/ncu_launches_$R.log 2>&1; echo "ncu step exit $?"
python tools/traffic_from_ncu.py gpurun_out/launches_${R}_cfg3.csv gpurun_out/${R}_traffic_cfg3.json
for W in decode encode t5; do
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_${R}_$W.csv \
   python tools/one_vae.py --what $W > gpurun_out/ncu_$W_$R.log 2>&1; echo "ncu $W exit $?"
done
