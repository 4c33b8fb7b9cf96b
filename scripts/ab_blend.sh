# development aid: forward/backward raster parity + A/B timing of blend variants (tagged libraries built beforehand)
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_raster_forward_gpu.py tests/test_raster_backward_gpu.py tests/test_pose_align_gpu.py -m gpu -x -q 2>&1 | tail -5
for tag in "" ${AB_TAGS:-}; do
 S3R_LIB_TAG=$tag timeout 300 python -m pytest tests/test_raster_forward_gpu.py -m gpu -x -q 2>&1 | tail -1
 for st in 1 8; do
  S3R_LIB_TAG=$tag python bench.py --steps 400 --warmup 20 --no-cpu --no-standin --no-encoder --legs none --streams $st > gpurun_out/ab_tmp.json 2> gpurun_out/ab_err.log || tail -5 gpurun_out/ab_err.log
  echo -n "tag=[$tag] streams=$st: "; python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['value']), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()}, round(d['roofline']['frac'],4), round(d.get('roofline_pipelined',{}).get('frac',0),4))" gpurun_out/ab_tmp.json
 done
done
