# development aid: value at 1 and 8 streams + stage times for the default library (or S3R_LIB_TAG)
cd $GRAFT_REPO_ROOT
for st in 1 8; do
  python bench.py --steps 400 --warmup 20 --no-cpu --no-standin --no-encoder --legs none --streams $st > gpurun_out/ab_tmp.json 2> gpurun_out/ab_err.log || tail -5 gpurun_out/ab_err.log
  echo -n "tag=[$S3R_LIB_TAG] streams=$st: "; python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(round(d['value']), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()}, round(d['roofline']['frac'],4), round(d.get('roofline_pipelined',{}).get('frac',0),4))" gpurun_out/ab_tmp.json
done
