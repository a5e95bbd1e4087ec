#!/bin/bash
# GPU test tier only.  usage: gpurun --timeout 1500 -- 'bash tools/gpu_pytest.sh <tag> [pytest args]'
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-pt}; shift
O=gpurun_out/$TAG
mkdir -p "$O"
( time timeout 1400 python -X faulthandler -m pytest tests -m gpu -q -x --durations=12 "$@" ) > "$O/pytest_gpu.log" 2>&1
head -60 "$O/pytest_gpu.log"; echo ...; tail -25 "$O/pytest_gpu.log"
