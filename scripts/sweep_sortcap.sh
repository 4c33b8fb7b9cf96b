#!/bin/bash
# development aid: shared-memory capacity of the per-tile sort (keys per CTA) vs occupancy; A/B/A/B to see the noise
mkdir -p gpurun_out
for cap in ${CAPS:-4096 3072 4096 3072 3584}; do
  touch styl3r_b200/csrc/raster_sort.cu styl3r_b200/csrc/s3r_common.cuh
  S3R_NVCC_FLAGS="-DS3R_SORT_SMEM_CAP=$cap" python -m styl3r_b200.build >/dev/null 2>&1 || { echo "build failed $cap"; continue; }
  python bench.py --steps 1000 --warmup 20 --no-cpu --no-standin --no-encoder > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_err.log || { echo "bench failed: $cap"; tail -3 gpurun_out/sweep_err.log; continue; }
  echo -n "cap=$cap: "; python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d[\"value\"]), d[\"stage_ms\"])" gpurun_out/sweep_tmp.json
done
