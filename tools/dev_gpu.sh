python -m pytest tests/test_training_gpu.py -m gpu -x -q 2>&1 | tail -25
