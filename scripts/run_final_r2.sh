# round-2 final measurements on the GPU box (one B200): GPU test suite, bench lines, reference arm
set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | grep -v Warning | tail -6 | cut -c1-250
python bench.py > gpurun_out/r2_bench_cfg2.json 2> gpurun_out/r2_bench_cfg2.err
python bench.py --no-cpu-baseline --fused 0 --steps 30 > gpurun_out/r2_bench_cfg2_unfused.json 2> /dev/null
python bench.py --no-cpu-baseline --config cfg4 --steps 30 > gpurun_out/r2_bench_cfg4.json 2> /dev/null
python bench.py --no-cpu-baseline --config cfg1 --steps 30 > gpurun_out/r2_bench_cfg1.json 2> /dev/null
python - <<'PY'
import json
for n in ("cfg2","cfg2_unfused","cfg4","cfg1"):
    try:
        d=json.loads(open("gpurun_out/r2_bench_%s.json"%n).read().strip().splitlines()[-1])
        f=d.get("roofline_fused") or {}
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "pageable", round(d["e2e"]["from_pageable_numpy"]["value"] or 0), "launches/step", d["gpu_launches_per_step"], "agg frac", round(d["roofline"]["frac"],3), "gemm frac", d["roofline_gemm"] and round(d["roofline_gemm"]["frac"],3), "fused ms", f.get("avg_launch_ms"), "cpu", d.get("cpu_baseline",{}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
PY
