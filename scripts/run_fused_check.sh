# fused-kernel iteration loop on the GPU box: parity tests, then the bench with the fused kernel on
timeout 120 python -m pytest tests/test_fused_gpu.py -x -q 2>&1 | grep -v Warning | tail -8 | cut -c1-250
if [ "${PIPESTATUS[0]}" != "0" ]; then echo TESTS FAILED; exit 1; fi
timeout 200 python bench.py --no-cpu-baseline --steps 30 > gpurun_out/bench_fused_iter.json 2> gpurun_out/bench_fused_iter.err || tail -c 800 gpurun_out/bench_fused_iter.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_fused_iter.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches_per_step"])
f=d.get("roofline_fused")
print("fused ms", f and f["avg_launch_ms"], "replaces", f and f["replaces_ms"], "frac", f and f["frac"])
PY
