# development aid: cfg3 / pose_align / cfg2_full legs for the default library and one tagged variant
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_raster_forward_gpu.py tests/test_raster_backward_gpu.py tests/test_compat_gpu.py -m gpu -x -q 2>&1 | tail -2
for tag in "" ${AB_TAGS:-tile}; do
  S3R_LIB_TAG=$tag python bench.py --steps 300 --warmup 20 --no-cpu --no-standin --no-encoder --legs cfg3,pose_align,cfg2_full > gpurun_out/ab_tmp.json 2> gpurun_out/ab_err.log || tail -5 gpurun_out/ab_err.log
  echo -n "tag=[$tag]: "; python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d['value']), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})
for k in ('cfg3','pose_align','cfg2_full'):
    v=d.get(k,{}); print('  ',k,{kk:(round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('ms_per_pass','views_per_s','encoder_ms','decoder_ms','total_ms','ms_per_step','stages_ms')})
" gpurun_out/ab_tmp.json
done
