"""Per-SASS-instruction executed counts / stall samples of an ncu report (source page), top N + opcode histogram.
   python scripts/ncu_sass_hot.py report.ncu-rep [N]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:end]))) if r.get("Instructions Executed") and r["Instructions Executed"].isdigit()]
tot = sum(int(r["Instructions Executed"]) for r in rows)
samp = sum(int(r["# Samples"]) for r in rows)
print("total warp instructions %d, samples %d, SASS lines %d" % (tot, samp, len(rows)))
ops = collections.Counter(); ops_s = collections.Counter()
for r in rows:
    op = r["Source"].split()[0] if not r["Source"].lstrip().startswith("@") else r["Source"].split()[1]
    op = op.split(".")[0]
    ops[op] += int(r["Instructions Executed"]); ops_s[op] += int(r["# Samples"])
print("opcode histogram (executed %, samples %):")
for op, n in ops.most_common(25):
    print("  %-10s %5.1f%%  %5.1f%%" % (op, 100 * n / tot, 100 * ops_s[op] / max(samp, 1)))
print("top %d SASS lines by samples:" % N)
for i, r in sorted(enumerate(rows), key=lambda t: -int(t[1]["# Samples"]))[:N]:
    print("  #%4d %6s smp %9s exec  %s" % (i, r["# Samples"], r["Instructions Executed"], r["Source"].strip()[:90]))
