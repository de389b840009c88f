#!/bin/bash
# round 2, GPU call 30 (1 GPU, last seconds of the budget): the two host-side edits after call 29 (smem limit of the pipelined row kernel, RoPE tables in the step-graph key)
timeout 60 python -m pytest tests/test_pipeline_gpu.py tests/test_ops_gpu.py -m gpu -q -k "graph or rmsnorm" 2>&1 | tail -3
