"""Instruction mix + stall-sample share per SASS opcode for kernels of an .ncu-rep (source page).
usage: ncu_instmix.py report.ncu-rep <substring of the demangled kernel name> [top]"""
import collections
import csv
import subprocess
import sys


def sections(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    name, hdr, data = None, None, []
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Kernel Name":
            if name is not None:
                yield name, hdr, data
            name, hdr, data = r[1], None, []
        elif r and r[0] == "Address":
            hdr = r
        elif hdr and len(r) > 6 and r[0].startswith("0x"):
            data.append(r)
    if name is not None:
        yield name, hdr, data


def main(path, sub, top=24):
    seen = set()
    for name, hdr, data in sections(path):
        if sub not in name or name in seen:
            continue
        seen.add(name)
        i_src, i_ex, i_st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
        tot = sum(int(r[i_ex]) for r in data)
        tots = sum(int(r[i_st]) for r in data) or 1
        ops, st = collections.Counter(), collections.Counter()
        for r in data:
            toks = r[i_src].strip().split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            op = op.split(".")[0]
            ops[op] += int(r[i_ex])
            st[op] += int(r[i_st])
        print(f"{name[:90]}: {tot} warp instructions, {tots} stall samples")
        for op, n in ops.most_common(top):
            print(f"  {op:12s} {n:12d} {100 * n / tot:5.1f}%   stall samples {100 * st[op] / tots:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 24)
