# round-end evidence at N = 1: the reference arm, the product arm, then the launch list of a short product run under ncu
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_final4_bench_reference_arm.json 2> gpurun_out/r02_final4_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final4_bench_n1.json 2> gpurun_out/r02_final4_bench_n1.err
tail -c 600 gpurun_out/r02_final4_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_final4_raw.csv python bench.py --steps 3 --warmup 3 > /dev/null 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02_final4_raw.csv > gpurun_out/r02_final4_launches_bench_n1.csv; rm -f gpurun_out/r02_final4_raw.csv
python -c "
import json
d=json.loads(open('gpurun_out/r02_final4_bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['fp32_pipe']['frac'])
print(d.get('path_stages')); print(d.get('loop',{}).get('ms_per_step')); print(d['parity'])
"
