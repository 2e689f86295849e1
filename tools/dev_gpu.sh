mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eam_gpu.py tests/test_hessian_gpu.py tests/test_setfl.py -m gpu -x -q -s 2>&1 | grep "E/N\|passed\|failed\|Error\|error\|assert" | tail -20
