#!/bin/bash
# development aid: rebuild the blend kernel with different tunables on the GPU box and bench each
mkdir -p gpurun_out
for cfg in "-DBLEND_U=4 -DBLEND_MINB=4" "-DBLEND_U=4 -DBLEND_MINB=3" "-DBLEND_U=8 -DBLEND_MINB=3" "-DBLEND_U=8 -DBLEND_MINB=2"; do
  S3R_NVCC_FLAGS="$cfg" python -m styl3r_b200.build --force >/dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  timeout 300 python -m pytest tests/test_raster_forward_gpu.py -m gpu -q -x 2>&1 | tail -1
  for st in 1 8; do
    python bench.py --steps 200 --warmup 10 --no-cpu --streams $st > gpurun_out/sweep_tmp.json 2> gpurun_out/sweep_err.log || { echo "bench failed: $cfg $st"; tail -3 gpurun_out/sweep_err.log; continue; }
    echo -n "$cfg streams=$st: "; python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d[\"value\"]), d[\"stage_ms\"])" gpurun_out/sweep_tmp.json
  done
done
