python tools/one_rvq.py 64 75 8 512 96 2>&1 | tail -2
python tools/one_rvq.py 4 64 8 512 96 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_callers.py tests/test_gpu_parity.py -x -q -m gpu -k "residual_vq or token_kernels or codec_against_golden or benchmarked or stages_teacher" 2>&1 | tail -3
