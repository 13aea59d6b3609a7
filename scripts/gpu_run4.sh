cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2d
timeout 900 python scripts/gpu_parity_diag.py 256 15 2>&1 | tee gpurun_out/r2d/diag_271.log | tail -24
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -s 2>&1 | tail -30 | tee gpurun_out/r2d/pytest_r2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "ipa or sample_loop or ga_encoder or se3" 2>&1 | tail -12 | tee gpurun_out/r2d/pytest_r1sel.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2d/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2d/bench_cfg4.json; python -c "
import json;d=json.load(open('gpurun_out/r2d/bench_cfg4.json'));print('cfg4: value %.2f ms/step %.3f e2e %.2f wall %.3f sample_wall %.2f ipa %.3f ms frac %.3f with_packers %.3f edge %.3f ms'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['wall_s'],d['sample_wall']['value'],d['roofline']['avg_launch_ms'],d['roofline']['frac'],d['roofline']['with_packers']['frac'],d['roofline_edge_transition']['avg_launch_ms']))"
