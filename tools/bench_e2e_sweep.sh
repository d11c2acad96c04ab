#!/bin/bash
# end-to-end host-buffer call (cfg5) vs slice count, with and without the kernels: where the time above the PCIe floor goes
for s in 1 2 4 8 16 32; do
  for k in 0 1; do
    echo "slices=$s skip_kernel=$k: $(MAC_HOST_SLICES=$s MAC_HOST_SKIP_KERNEL=$k python tools/bench_e2e.py quick 2>&1 | tail -1)"
  done
done
