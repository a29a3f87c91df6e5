#!/bin/bash
# quick GPU regression of the VQ path (parity tests of the tensor-core / refine variants) + the config-2 A/B line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vq_gpu.py -x -q -m gpu ${PYTEST_K:+-k "$PYTEST_K"} 2>&1 | tail -15 | tee gpurun_out/quicktest.log
TAG=base timeout 120 python scripts/ab_config2.py 2>&1 | tail -2 | cut -c1-500
