#!/bin/bash
# GPU call 67: frames in parts: the new parity test, then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "in_parts" 2>&1 | tail -12 ) 2>&1 | tee gpurun_out/r04m_parts.log
( timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r04m_pytest.log
