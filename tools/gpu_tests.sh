#!/bin/bash
# all GPU tests (what the driver runs at round end) + a log
export TAG=${1:-r2t}
python -m pytest tests -x -q -m gpu --durations=15 > gpurun_out/${TAG}_pytest.log 2>&1
tail -25 gpurun_out/${TAG}_pytest.log
