#!/bin/bash
# A/B on one box: the tree before this session (5fa3abf, built under _old_tree/) vs HEAD on cfg4's aggregation
mkdir -p gpurun_out
(cd _old_tree && timeout 300 python bench.py --config cfg4 --steps 3 --no-cpu > ../gpurun_out/r2c19_cfg4_old.json 2> ../gpurun_out/r2c19_cfg4_old.err)
timeout 300 python bench.py --config cfg4 --steps 3 --no-cpu > gpurun_out/r2c19_cfg4_new.json 2> gpurun_out/r2c19_cfg4_new.err
python - <<'PY'
import json
for k in ("old", "new"):
    try:
        d = json.loads(open("gpurun_out/r2c19_cfg4_%s.json" % k).read().strip().splitlines()[-1])
        print(k, "ms/step %.2f" % d["ms_per_step"], [round(l["aggregation_ms"], 2) for l in d["config"]["layers"]])
    except Exception as ex:
        print(k, "unparsed", ex, open("gpurun_out/r2c19_cfg4_%s.err" % k).read()[-400:])
PY
