#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2j}
( time timeout 900 python -m pytest tests -m gpu -q -k "next_scope or refused" ) > gpurun_out/${TAG}_next_pytest.log 2>&1
echo "next pytest exit $?" >> gpurun_out/${TAG}_next_pytest.log
tail -40 gpurun_out/${TAG}_next_pytest.log
