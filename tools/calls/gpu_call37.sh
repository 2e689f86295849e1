#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_domain_gpu.py -m gpu -q -x 2>&1 | tail -5
