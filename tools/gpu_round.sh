#!/bin/bash
# usage: tools/gpu_round.sh TAG  -- tests + smoke + quick timings on the GPU box
TAG=${1:-rX}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/${TAG}_smoke.log
timeout 300 python tools/quick_time.py > gpurun_out/${TAG}_quick.log 2>&1; echo "quick exit $?" >> gpurun_out/${TAG}_quick.log
tail -n 3 gpurun_out/${TAG}_pytest.log; tail -n 2 gpurun_out/${TAG}_smoke.log; tail -n 12 gpurun_out/${TAG}_quick.log
