#!/bin/bash
mkdir -p gpurun_out
for cfg in "-DBLEND_CORNER_CUT=0 -DBLEND_SKIP_DEAD=0" "-DBLEND_CORNER_CUT=0 -DBLEND_SKIP_DEAD=1" "-DBLEND_CORNER_CUT=1 -DBLEND_SKIP_DEAD=0"; do
  S3R_NVCC_FLAGS="$cfg" python -m styl3r_b200.build --force >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  for st in 1 8; do
    python bench.py --steps 300 --warmup 10 --no-cpu --streams $st > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_err.log || { echo "bench failed: $cfg $st"; tail -3 gpurun_out/sweep_err.log; continue; }
    echo -n "$cfg streams=$st: "; python scripts/pj.py gpurun_out/sweep_tmp.json
  done
done
