ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02ad_occ_raw.csv python tools/bench_nets.py --iters 1 --chunk 65536 > gpurun_out/r02ad_occ.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/r02ad_occ_raw.csv > gpurun_out/r02ad_occ_launches.csv
rm -f gpurun_out/r02ad_occ_raw.csv
grep "^#" gpurun_out/r02ad_occ_launches.csv | head -30
tail -1 gpurun_out/r02ad_occ.log | cut -c1-600
