# kernel-only timing of the coverage-gain instruction-mix variants (MAC_COVGAIN_VARIANT), 512 and 64 cameras
for v in ${VARIANTS:-0 2 7}; do MAC_COVGAIN_VARIANT=$v python tools/bench_covgain.py 2 ; done
