set -x
python bench.py --no-cpu-baseline > gpurun_out/bench_r2g_fused.json 2> gpurun_out/bench_r2g_fused.err
python bench.py --no-cpu-baseline --fused 0 --steps 20 > gpurun_out/bench_r2g_unfused.json 2> gpurun_out/bench_r2g_unfused.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_layer_kernel -s 3 -c 1 -o gpurun_out/fused_layer -f python bench.py --steps 1 --warmup 3 --cuda-graph 0 --no-cpu-baseline > gpurun_out/ncu_fused.log 2>&1
tail -c 1500 gpurun_out/bench_r2g_fused.err
python - <<'PY'
import json
for n in ("fused","unfused"):
    try:
        d=json.loads(open("gpurun_out/bench_r2g_%s.json"%n).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "launches/step", d["gpu_launches_per_step"])
        print("  roofline", d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], "gemm", d["roofline_gemm"]["frac"] if d["roofline_gemm"] else None, d["roofline_gemm"]["avg_launch_ms"] if d["roofline_gemm"] else None)
        print("  fused", d.get("roofline_fused"))
        print("  shares", {k: (round(v,3) if isinstance(v,float) else v) for k,v in d["shares"].items() if k!="_note"})
    except Exception as e:
        print(n, "ERR", e)
PY
