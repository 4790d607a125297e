"""SASS of one kernel in address order with executed counts and stall samples, from an ncu report.
usage: python profiles/ncu_sass.py REPORT.ncu-rep kernel-substring > out.txt"""
import csv
import io
import subprocess
import sys

rep, want = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
use, header = False, None
out = {}
for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] == "Function Name":
        use = want in r[1]
        continue
    if r[0] == "Line No":
        header = r
        i_exec, i_smp = header.index("Instructions Executed"), header.index("# Samples")
        continue
    if r[0] == "File Path" or not use or header is None:
        continue
    if r[0] == "":
        cur_line = last_line
        try:
            out[int(r[2], 16)] = (r[3].strip(), int(r[i_exec]), int(r[i_smp]), cur_line)
        except ValueError:
            pass
    else:
        last_line = r[0]
base = min(out)
for a in sorted(out):
    t, n, s, ln = out[a]
    print(f"{a-base:6x} {n:10d} {s:5d}  L{ln:>4s}  {t}")
