cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2f
T="timeout 600 python -m pepflowww_b200.train --out gpurun_out/r2f/train_1gpu.jsonl"
$T --iters 8 --warmup 3 --batch-size 32 --pocket 48 --peptide 12 --graph 2>&1 | tail -4
$T --iters 5 --warmup 2 --batch-size 32 --pocket 128 --peptide 12 --graph 2>&1 | tail -3
$T --iters 5 --warmup 2 --batch-size 32 --pocket 128 --peptide 12 --graph --tf32 2>&1 | tail -3
$T --iters 5 --warmup 2 --batch-size 32 --pocket 128 --peptide 12 --tf32 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -s -k "200_step" 2>&1 | tail -8
