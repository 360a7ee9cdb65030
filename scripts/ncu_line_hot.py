"""Per-CUDA-source-line executed warp instructions / stall samples of an ncu report.
   python scripts/ncu_line_hot.py report.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Line No"'))
rd = csv.reader(io.StringIO("\n".join(lines[start:])))
hdr = next(rd)
iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
rows = [(int(r[0]), r[1], int(r[iE]), int(r[iS]) if r[iS].isdigit() else 0) for r in rd if r and r[0].isdigit() and r[iE].isdigit()]
tot = sum(r[2] for r in rows); samp = sum(r[3] for r in rows)
print("total %d warp instr, %d samples" % (tot, samp))
for ln, src, e, s in sorted(rows, key=lambda r: -r[2])[:N]:
    print("%5d %5.1f%% exec %5.1f%% smp  %s" % (ln, 100 * e / tot, 100 * s / max(samp, 1), src.strip()[:110]))
