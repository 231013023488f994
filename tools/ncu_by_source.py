"""Dynamic warp-instructions and stall samples of one kernel of an .ncu-rep, aggregated by SOURCE REGION of the calling file
(default llq_step16.cuh; LLQ_SRC_FILE overrides): SASS addresses from the report's source page are mapped to `//## File ... line N`
markers of `nvdisasm -g` on the same library (inlined helpers from other headers are attributed to the last line of the calling file
seen before them).
Usage: python tools/ncu_by_source.py REP LIB.so KERNEL_MANGLED_SUBSTR [bucket]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l)
    end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith(".text.")), len(dis))
    addr2line, klast = {}, None
    for l in dis[start:end]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            if m.group(1).endswith(os.environ.get("LLQ_SRC_FILE", "llq_step16.cuh")):
                klast = int(m.group(2))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/", l)
        if m:
            addr2line[int(m.group(1), 16)] = klast
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(rows[hi])}
    base = None
    samp, inst = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) < len(col):
            continue
        a = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
        if base is None:
            base = a
        ln = addr2line.get(a - base)
        b = None if ln is None else ln // bucket * bucket
        samp[b] += int(r[col["# Samples"]] or 0)
        inst[b] += int(r[col["Instructions Executed"]] or 0)
    ts, ti = sum(samp.values()) or 1, sum(inst.values()) or 1
    print("# %s: %d warp-instructions, %d samples; rows = %s lines [b, b+%d)" % (os.path.basename(rep), ti, ts, os.environ.get("LLQ_SRC_FILE", "llq_step16.cuh"), bucket))
    for b in sorted(samp, key=lambda x: (x is None, x)):
        print("%6s  inst %5.1f%%  samples %5.1f%%" % (b, 100.0 * inst[b] / ti, 100.0 * samp[b] / ts))


if __name__ == "__main__":
    main()
