# N = 2 check after the kernel changes of the round: the multi-GPU tests, then the bench line (stages skipped)
timeout 600 python -m pytest tests/test_multigpu.py tests/test_covgain_gpu.py -q -m gpu > gpurun_out/r02_final2_pytest_2gpu.log 2>&1; tail -2 gpurun_out/r02_final2_pytest_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-stages --no-cpu-baseline > gpurun_out/r02_final2_bench_n2.json 2> gpurun_out/r02_final2_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_final2_bench_n2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['parity'], d.get('loop',{}).get('ms_per_step'))
"
