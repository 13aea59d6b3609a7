cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r2c
timeout 300 python scripts/gpu_gemm_error.py 2>&1 | tee gpurun_out/r2c/gemm_error.log | tail -14
timeout 900 python scripts/gpu_parity_diag.py 256 15 2>&1 | tee gpurun_out/r2c/diag_271.log | tail -24
timeout 900 python -m pytest tests/test_gpu_parity_r2.py -q -s 2>&1 | tail -40 | tee gpurun_out/r2c/pytest_r2.log
