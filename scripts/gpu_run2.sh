cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2b
timeout 900 python scripts/gpu_parity_diag.py 256 15 2>&1 | tee gpurun_out/r2b/diag_271.log | tail -30
timeout 600 python scripts/gpu_parity_diag.py 24 6 2>&1 | tee gpurun_out/r2b/diag_30.log | tail -30
for t in 0 1 3; do timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --edge-terms $t 2>&1 | tail -1 > gpurun_out/r2b/bench_terms$t.json; python -c "
import json;d=json.load(open('gpurun_out/r2b/bench_terms$t.json'));print('edge_terms $t: ms/step %.3f edge ms %.3f frac %.3f'%(d['ms_per_step'],d['roofline_edge_transition']['avg_launch_ms'],d['roofline_edge_transition']['frac']))"; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r2b/bench_cfg4.json; python -c "
import json;d=json.load(open('gpurun_out/r2b/bench_cfg4.json'));print('cfg4: value %.2f e2e %.2f wall %.3f sample_wall %.2f'%(d['value'],d['e2e']['value'],d['e2e']['wall_s'],d['sample_wall']['value']))"
