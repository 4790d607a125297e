"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report (needs -lineinfo).
usage: python profiles/ncu_lines.py REPORT.ncu-rep [kernel-substring] [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, cur_fn, header, use = None, None, None, False
agg = {}
total_i = total_s = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        use = want in cur_fn
        continue
    if r[0] == "Line No":
        header = r
        i_exec = header.index("Instructions Executed")
        i_smp = header.index("# Samples")
        continue
    if not use or header is None or r[0] == "":
        continue                      # SASS rows (empty line number) are already summed in their source row
    try:
        n, s = int(r[i_exec]), int(r[i_smp])
    except ValueError:
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:90])
    a = agg.setdefault(key, [0, 0])
    a[0] += n
    a[1] += s
    total_i += n
    total_s += s
print(f"total {total_i} warp instructions, {total_s} samples")
for (f, ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/total_i:5.1f}% inst {100*s/max(total_s,1):5.1f}% smp  {f}:{ln:<4d} {src}")
