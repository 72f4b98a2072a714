"""Hot instructions of one kernel in an ncu report (needs -lineinfo + --import-source on):
    python tools/ncu_source_hot.py <report.ncu-rep> <kernel regex> [top N]
Prints the SASS lines with the most warp-stall samples, their stall reasons and shared-memory wavefront excess, plus an opcode histogram."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    hi = next(i for i, x in enumerate(r) if x and x[0] == "Address")
    h = r[hi]
    rows = [x for x in r[hi + 1:] if len(x) == len(h) and x[0].startswith("0x")]
    col = {n: i for i, n in enumerate(h)}
    S = lambda x, n: int(float(x[col[n]] or 0))
    tot = sum(S(x, "# Samples") for x in rows)
    totex = sum(S(x, "Instructions Executed") for x in rows)
    print(f"{len(rows)} SASS lines, {tot} samples, {totex} warp instructions")
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = collections.Counter()
    for x in rows:
        for n in stalls:
            agg[n] += S(x, n)
    print("stall totals:", [(k, v) for k, v in agg.most_common(8)])
    for x in sorted(rows, key=lambda x: -S(x, "# Samples"))[:topn]:
        why = sorted(((S(x, n), n[6:]) for n in stalls), reverse=True)[:2]
        print(f"{S(x, '# Samples'):6d} {S(x, '# Samples') / max(tot, 1):6.3f} ex={S(x, 'Instructions Executed'):9d} shw={x[col['L1 Wavefronts Shared']]:>9s}/{x[col['L1 Wavefronts Shared Ideal']]:>9s} "
              f"{why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}  {x[col['Source']].strip()[:80]}")
    hist = collections.Counter()
    for x in rows:
        t = x[col["Source"]].strip().split()
        op = t[1] if t and t[0].startswith("@") else (t[0] if t else "?")
        hist[op.split(".")[0]] += S(x, "Instructions Executed")
    print("opcodes:", [(k, round(v / max(totex, 1), 3)) for k, v in hist.most_common(24)])


if __name__ == "__main__":
    main()
