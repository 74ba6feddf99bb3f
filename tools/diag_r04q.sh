#!/bin/bash
# GPU call 72: the whole GPU suite on the final build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1 | tee gpurun_out/r04q_pytest.log
