#!/bin/bash
# full GPU test suite + the two bench lines (no extras) -> gpurun_out/${TAG}_*
mkdir -p gpurun_out
TAG=${1:-chk}
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --no-extras --no-eager --no-cpu-baseline > gpurun_out/${TAG}_train.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload rollout --no-eager --no-cpu-baseline > gpurun_out/${TAG}_rollout.json 2>> gpurun_out/${TAG}_bench.err
tail -4 gpurun_out/${TAG}_pytest.log
python - <<PY
import json
for f in ("train","rollout"):
    try:
        d=json.loads(open(f"gpurun_out/${TAG}_{f}.json").read().strip().splitlines()[-1])
        r=d.get("roofline",{})
        print(f, round(d["value"],1), d["unit"], "ms/step", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), "all_gemm_frac", round(r.get("all_gemm_frac_of_tensor_peak",0),3), {k:(round(v["frac"],3), round(v["avg_us_per_launch"],1)) for k,v in r.get("classes",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
