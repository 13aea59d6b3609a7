set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2a
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -s -x 2>&1 | tail -60 > gpurun_out/r2a/pytest_r2.log
cat gpurun_out/r2a/pytest_r2.log | tail -40
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "sample_loop or embed or euler or launch_counter or flags" 2>&1 | tail -15 > gpurun_out/r2a/pytest_r1sel.log
tail -8 gpurun_out/r2a/pytest_r1sel.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/r2a/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a/bench_cfg4.json 2> gpurun_out/r2a/bench_cfg4.err; tail -c 3000 gpurun_out/r2a/bench_cfg4.json; tail -5 gpurun_out/r2a/bench_cfg4.err
timeout 300 python bench.py --config cfg1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_cfg1.json 2> gpurun_out/r2a/bench_cfg1.err; tail -c 1500 gpurun_out/r2a/bench_cfg1.json; tail -5 gpurun_out/r2a/bench_cfg1.err
timeout 300 python bench.py --config cfg1 --graph off --steps 50 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2a/bench_cfg1_nograph.json 2>&1; tail -c 600 gpurun_out/r2a/bench_cfg1_nograph.json
timeout 300 python bench.py --config cfg3 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2a/bench_cfg3.json 2> gpurun_out/r2a/bench_cfg3.err; tail -c 1500 gpurun_out/r2a/bench_cfg3.json
