#!/bin/bash
# development aid: blend kernel CTA granularity (CTAs per 16x16 tile) - rebuild on the GPU box and bench each
mkdir -p gpurun_out
for cfg in "-DBLEND_SPLIT=1 -DBLEND_MINB=3" "-DBLEND_SPLIT=2 -DBLEND_MINB=6" "-DBLEND_SPLIT=4 -DBLEND_MINB=8" "-DBLEND_SPLIT=8 -DBLEND_MINB=8" "-DBLEND_SPLIT=4 -DBLEND_MINB=8 -DBLEND_STAGES=3" "-DBLEND_SPLIT=1 -DBLEND_MINB=3"; do
  touch styl3r_b200/csrc/raster_blend.cu
  S3R_NVCC_FLAGS="$cfg" python -m styl3r_b200.build >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  timeout 300 python -m pytest tests/test_raster_forward_gpu.py -m gpu -q -x 2>&1 | tail -1
  for st in 1 8; do
    python bench.py --steps 300 --warmup 10 --no-cpu --no-standin --no-encoder --streams $st > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_err.log || { echo "bench failed: $cfg $st"; tail -3 gpurun_out/sweep_err.log; continue; }
    echo -n "$cfg streams=$st: "; python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d[\"value\"]), d[\"stage_ms\"])" gpurun_out/sweep_tmp.json
  done
done
