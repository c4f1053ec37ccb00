#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_faces.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tee gpurun_out/pytest_faces.log | tail -40
