# round-end check on the GPU box: full GPU test suite, smoke, the default bench line, the reference arm, launch list
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -2 gpurun_out/final_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-standin --no-encoder --legs none > /dev/null 2>&1; wc -l gpurun_out/final_launches.csv
