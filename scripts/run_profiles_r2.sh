# round-2 profile captures (one B200): launch list of the eager step + full capture of the fused layer kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 520 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 1 --warmup 3 --cuda-graph 0 --no-cpu-baseline > gpurun_out/launches_r2b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused_layer_kernel -s 3 -c 1 -o gpurun_out/fused_layer_final -f python bench.py --steps 1 --warmup 3 --cuda-graph 0 --no-cpu-baseline > gpurun_out/ncu_fused_final.log 2>&1
tail -2 gpurun_out/ncu_fused_final.log; wc -l gpurun_out/launches_r2b.csv
