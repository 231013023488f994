"""Text summary of one kernel of an .ncu-rep (raw metrics + stall sampling totals + hottest SASS lines by source line).
Usage: python tools/ncu_summary.py gpurun_out/step_v7.ncu-rep [n_hot]"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
           "launch__shared_mem_per_block_static", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    n_hot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = page(rep, "raw")
    d = dict(zip(raw[0], zip(raw[2], raw[1])))
    print("# kernel:", d.get("Kernel Name", ("?",))[0][:120])
    for m in METRICS:
        if m in d:
            print("%-85s %s %s" % (m, d[m][0], d[m][1]))
    src = page(rep, "source")
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    hdr = src[hi]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    lines = []
    for r in src[hi + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            ns = int(r[col["# Samples"]] or 0)
            ni = int(r[col["Instructions Executed"]] or 0)
        except ValueError:
            continue
        for h in stall_cols:
            try:
                tot[h] += int(r[col[h]] or 0)
            except ValueError:
                pass
        lines.append((ns, ni, r[col["Source"]]))
    total = sum(tot.values()) or 1
    print("\n# warp stall sampling (all samples), share of samples")
    for h, v in tot.most_common(9):
        print("%-28s %5.1f%%" % (h, 100.0 * v / total))
    tot_i = sum(x[1] for x in lines) or 1
    tot_s = sum(x[0] for x in lines) or 1
    print("\n# hottest SASS instructions by stall samples (share of samples, share of executed warp-instructions)")
    for ns, ni, s in sorted(lines, reverse=True)[:n_hot]:
        print("%5.2f%% %5.2f%%  %s" % (100.0 * ns / tot_s, 100.0 * ni / tot_i, re.sub(r"\s+", " ", s)[:110]))
    ops = collections.Counter()
    for ns, ni, s in lines:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
        if m:
            ops[m.group(2)] += ni
    print("\n# executed warp-instructions by opcode")
    print(", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in ops.most_common(16)))


if __name__ == "__main__":
    main()
