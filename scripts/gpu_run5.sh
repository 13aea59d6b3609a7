cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2e
T="timeout 600 python -m pepflowww_b200.train --out gpurun_out/r2e/train_1gpu.jsonl"
$T --iters 8 --warmup 3 --batch-size 32 --pocket 48 --peptide 12 2>&1 | tail -2
$T --iters 8 --warmup 3 --batch-size 32 --pocket 48 --peptide 12 --graph 2>&1 | tail -3
$T --iters 5 --warmup 2 --batch-size 32 --pocket 128 --peptide 12 2>&1 | tail -2
$T --iters 5 --warmup 2 --batch-size 32 --pocket 128 --peptide 12 --graph 2>&1 | tail -3
$T --iters 5 --warmup 2 --batch-size 32 --pocket 128 --peptide 12 --graph --tf32 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "training" 2>&1 | tail -3
