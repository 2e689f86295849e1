mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_atomic_gpu.py tests/test_training_gpu.py -m gpu -x -q -s 2>&1 | grep "dG\|passed\|failed\|Error\|assert" | tail -20
