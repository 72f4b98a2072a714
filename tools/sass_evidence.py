"""Lists, per kernel of relion_b200/librelion_b200.so, the SASS mnemonics that show what the kernel is built on
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, cp.async -> LDGSTS, red.global -> REDG,
cluster barrier -> UCGABAR, 256-bit global loads -> LDG.E.ENL2.256).  Works without a GPU:  python tools/sass_evidence.py > profiles/sass_evidence_rNN.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r'\b(UTC[A-Z]*MMA(?:\.[A-Z0-9_.]+)?|UTMALDG(?:\.[A-Z0-9_.]+)?|UTCBAR(?:\.[A-Z0-9_.]+)?|UTCATOMSWS(?:\.[A-Z0-9_.]+)?|'
                 r'LDTM(?:\.[A-Z0-9_.x]+)?|LDG\.E\.ENL2\.256(?:\.[A-Z0-9_.]+)?|LDGSTS(?:\.[A-Z0-9_.]+)?|REDG(?:\.[A-Za-z0-9_.]+)?|UCGABAR_[A-Z]+|ATOMG(?:\.[A-Z0-9_.]+)?)')


def main():
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "relion_b200", "librelion_b200.so")],
                          capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            for x in PAT.findall(line):
                cnt[cur][x] += 1
    print("# cuobjdump -sass relion_b200/librelion_b200.so (sm_100a), tools/sass_evidence.py: per kernel, counts of the mnemonics that")
    print("# show tcgen05 MMAs (UTC*MMA, .2CTA = cta_group::2), TMEM loads (LDTM), TMA (UTMALDG), cp.async (LDGSTS), vector")
    print("# reductions (REDG ... F32x4 = red.global.add.v4.f32), cluster barriers (UCGABAR) and 256-bit global loads (LDG.E.ENL2.256 = ld.global.nc.v8.f32)")
    for f, c in cnt.items():
        if any(k.startswith(("UTC", "UTMA", "LDTM", "LDGSTS", "REDG.E.ADD.F32")) for k in c):
            name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip().split("(")[0]
            print(f"{name}: " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))


if __name__ == "__main__":
    main()
