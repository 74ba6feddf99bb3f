#!/bin/bash
# round 2, GPU call 14: whole GPU suite (no -x), then the ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 ) > gpurun_out/r02n_pytest.log
bash tools/profile_r02.sh > gpurun_out/r02n_profile.log 2>&1
echo done
