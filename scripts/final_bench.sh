# round-end bench lines (default arm + reference arm) into gpurun_out/
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; tail -2 gpurun_out/final_bench_ref.err
