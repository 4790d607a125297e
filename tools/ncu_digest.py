"""Digest of one ncu report (first kernel): headline metrics, stall reasons, dynamic instruction mix and the
instruction segments between synchronisation points.  Usage: python tools/ncu_digest.py REPORT.ncu-rep [--kernel REGEX] [--segments] [--hot N]"""
import collections
import csv
import io
import subprocess
import sys


def page(rep, name, extra=()):
    extra = (*extra, "-c", "1")                  # first launch (whose name matches --kernel)
    if "--kernel" in sys.argv:
        extra = (*extra, "-k", "regex:" + sys.argv[sys.argv.index("--kernel") + 1])
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.per_cycle_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    h, u, v = raw[0], raw[1], raw[2]
    d = dict(zip(h, zip(u, v)))
    print("kernel:", d.get("Kernel Name", ("", ""))[1][:120])
    for k in KEYS:
        if k in d:
            print(f"  {k:75s} {d[k][1]:>16s} {d[k][0]}")
    src = page(rep, "source", ["--print-source", "sass"])
    hdr, data = src[1], src[2:]
    for k, r in enumerate(data):                 # a report with several launches repeats the header per launch: keep the first
        if len(r) < 10:
            data = data[:k]
            break
    iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    st = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[iE]) for r in data)
    tots = sum(int(r[iSm]) for r in data) or 1

    def op(r):
        s = r[iS].split()
        return (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    stall = collections.Counter()
    ops = collections.Counter()
    smp = collections.Counter()
    for r in data:
        for i, n in st:
            stall[n] += int(r[i])
        ops[op(r)] += int(r[iE])
        smp[op(r)] += int(r[iSm])
    print(f"warp instructions {tot}, samples {tots}")
    ssum = sum(stall.values()) or 1
    print("stalls: " + ", ".join(f"{n[6:]} {100 * c / ssum:.1f}%" for n, c in stall.most_common(9)))
    print("mix:    " + ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in ops.most_common(24)))
    if "--segments" in sys.argv:
        seg, cur, cure, curs, start = [], collections.Counter(), 0, 0, 0
        for k, r in enumerate(data):
            o = r[iS].split()
            o = o[1] if o[0].startswith("@") else o[0]
            cur[o.split(".")[0]] += int(r[iE]); cure += int(r[iE]); curs += int(r[iSm])
            if o.startswith("SYNCS") or o.startswith("BAR"):
                seg.append((start, k, cure, curs, cur, r[iS].strip()[:60])); cur = collections.Counter(); cure = curs = 0; start = k + 1
        seg.append((start, len(data), cure, curs, cur, "end"))
        for s in seg:
            if s[2] > tot * 0.004 or s[3] > tots * 0.004:
                print(f"  [{s[0]:5d},{s[1]:5d}] {100 * s[2] / tot:5.1f}% instr {100 * s[3] / tots:5.1f}% smp  {s[5]:60s} {dict(s[4].most_common(6))}")
    if "--hot" in sys.argv:
        n = int(sys.argv[sys.argv.index("--hot") + 1])
        rows = sorted(enumerate(data), key=lambda kr: -int(kr[1][iSm]))[:n]
        for k, r in sorted(rows):
            top = sorted(((int(r[i]), nm[6:]) for i, nm in st), reverse=True)[:2]
            print(f"  {k:5d} exec {int(r[iE]):9d} smp {int(r[iSm]):5d}  {r[iS].strip()[:80]:80s} {top}")


if __name__ == "__main__":
    main()
