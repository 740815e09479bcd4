#!/usr/bin/env bash
# compute-sanitizer targets for the hand-written kernels (the reference has no race tooling; every kernel here
# that synchronises through shared memory, mbarriers, tickets or peer flags gets a memcheck / racecheck / synccheck pass).
# Usage on a GPU box:  tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
set -euo pipefail
TOOL=${1:-memcheck}
EXPR=${2:-"gemv or attention or rope or argmax or sample or moe or quantized or gemm"}
export NXDI_B200_PDL=0            # programmatic dependent launch confuses racecheck's per-kernel hazard tracking
export CUDA_LAUNCH_BLOCKING=1
exec compute-sanitizer --tool "$TOOL" --error-exitcode 99 --launch-timeout 120 \
    python -m pytest tests/test_kernels_gpu.py -x -q -k "$EXPR"
