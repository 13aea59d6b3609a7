# The commands of the 8 x B200 call of round 2 (sharded-sample equality test, cfg5 DDP training at 2 / 4 / 8 GPUs, eager with
# the NCCL overlap profile and with the whole-iteration CUDA graph).  Results: profiles/r2_train_cfg5_ddp.jsonl,
# profiles/r2_pytest_sharded_2gpu.log.  Run under: gpurun --gpus 8 -- bash scripts/gpu_cfg5_ddp_runs.sh
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2g
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_headline_parity.py -q -s -k "sharded" 2>&1 | tail -4 | tee gpurun_out/r2g/pytest_sharded.log
run() { N=$1; shift; timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) -m pepflowww_b200.train --batch-size 32 --pocket 128 --peptide 12 --out gpurun_out/r2g/train_cfg5.jsonl "$@" 2>&1 | grep -v "^W\|warn\|^$" | tail -3; }
run 2 --iters 6 --warmup 2 --profile 2
run 2 --iters 6 --warmup 2 --graph
run 4 --iters 6 --warmup 2 --profile 2
run 8 --iters 6 --warmup 2 --profile 2
run 8 --iters 6 --warmup 2 --graph
run 8 --iters 6 --warmup 2 --graph --tf32
cat gpurun_out/r2g/train_cfg5.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); a = d.get('ddp_allreduce') or {}
    print('N=%d graph=%s tf32=%s  %.1f ms/iter  %.1f samples/s  nccl %s ms/iter overlap %s' % (d['n_gpus'], d['cuda_graph'], d['tf32_matmul'], d['ms_per_iter'], d['value'], a.get('nccl_kernel_ms_per_iter'), a.get('overlapped_with_compute')))"
