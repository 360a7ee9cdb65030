"""Turn gpurun_out/*.ncu-rep / launch lists into the small text summaries committed under profiles/."""
import collections, csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_st.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "smsp__sass_inst_executed_op_utcmma.sum", "smsp__sass_inst_executed_op_tma_ld.sum",
        "smsp__sass_inst_executed_op_tma_st.sum",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_not_selected",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def full_summary(rep, title):
    hdr, units, data = raw(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["# %s" % title, "# source: %s (ncu --set full --clock-control none --import-source on)" % rep, ""]
    names = [d[idx["Kernel Name"]] for d in data]
    lines.append("%-92s %s" % ("metric [unit]", " | ".join("launch %d" % i for i in range(len(data)))))
    lines.append("%-92s %s" % ("kernel", " | ".join(n.replace("void gr::<unnamed>::", "")[:40] for n in names)))
    for k in KEYS:
        if k in idx:
            lines.append("%-92s %s" % ("%s [%s]" % (k, units[idx[k]]), " | ".join(d[idx[k]] for d in data)))
    return "\n".join(lines) + "\n", [dict((k, d[idx[k]]) for k in KEYS if k in idx) for d in data]


def launch_summary(csv_path, nfwd):
    with open(csv_path) as f:
        rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = r["Kernel Name"]
        key = ("ours: " + n.split("(")[0].replace("void ", "").replace("gr::<unnamed>::", "")) if "gr::" in n \
            else ("torch: " + n.split("(")[0].replace("void ", "")[:60])
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        agg[key][0] += 1
        agg[key][1] += v
    nfwd = max(1, sum(1 for r in rows if "hist_kernel" in r["Kernel Name"]))   # one CSR build per forward
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("ours"))
    lines = ["# per-kernel device time, %d forwards (ncu --metrics gpu__time_duration.sum --clock-control none;" % nfwd,
             "# cold-cache, serialised: compare SHARES, not absolutes).  source: %s" % csv_path,
             "# total %.0f us over %d launches = %.0f us / forward; our kernels %.1f%% of device time" % (
                 tot, len(rows), tot / nfwd, 100 * ours / tot), ""]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        lines.append("%-75s n=%4d total=%9.1f us avg=%8.1f us %5.1f%%" % (k[:75], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    return "\n".join(lines) + "\n"


if __name__ == "__main__":
    tag = sys.argv[1]
    open("profiles/%s_launches.txt" % tag, "w").write(launch_summary("gpurun_out/launches_%s.csv" % tag, 5))
    t, d = full_summary("gpurun_out/agg_%s.ncu-rep" % tag, "aggregation kernel (gr_aggregate_dual_abs, persistent warp-specialised), two dense-prior launches (layers 1 and 2 of iteration 0); cfg2")
    open("profiles/%s_agg_kernel.txt" % tag, "w").write(t)
    traffic = [(float(x["dram__bytes_read.sum"]) + float(x["dram__bytes_write.sum"])) * 1e6 for x in d]
    t2, d2 = full_summary("gpurun_out/tc_%s.ncu-rep" % tag, "tcgen05 e2e GEMM (gr_linear_tc_planes), cfg2: M=128000 N=200 K=1040")
    open("profiles/%s_tc_gemm.txt" % tag, "w").write(t2)
    json.dump({"agg_dense_traffic_bytes_per_launch": traffic[-1],
               "agg_dense_ncu_us": float(d[-1]["gpu__time_duration.sum"]), "source": "profiles/%s_agg_kernel.txt" % tag},
              open("profiles/%s_traffic.json" % tag, "w"), indent=1)
    print(open("profiles/%s_launches.txt" % tag).read()[:2500]); print(t); print(t2)
