# depth network: parity tests, forward time (implicit GEMM + split-K | no split-K | explicit im2col), per-launch device times
timeout 600 python -m pytest tests/test_depth_gpu.py tests/test_macarons_step_gpu.py -x -q 2>&1 | tail -3
python tools/bench_nets.py --skip-occ --depth --iters 20 | tail -1
MAC_DEPTH_GRAPH=0 python tools/bench_nets.py --skip-occ --depth --iters 20 | tail -1
MAC_DEPTH_IM2COL=1 python tools/bench_nets.py --skip-occ --depth --iters 20 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear_kernel|im2col|cost_volume|maxpool|nchw|splitk" --csv --log-file gpurun_out/r02u_depth_raw.csv python tools/bench_nets.py --skip-occ --depth --iters 1 > /dev/null 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02u_depth_raw.csv > gpurun_out/r02u_depth_launches.csv
rm -f gpurun_out/r02u_depth_raw.csv
