"""Blackwell-native evidence: per-kernel counts of the SASS mnemonics that the tcgen05 / TMEM / TMA / cluster PTX compiles to
(UTCHMMA / UTCQMMA ... = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor, UCGABAR /
BAR.SYNC on cluster scope = barrier.cluster, FFMA2 = fma.rn.f32x2) in the shipped library.
    python profiles/sass_summary.py > profiles/sass_summary.txt      (cuobjdump only; no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpzoo_b200", "lib", "libgpzoo_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UCGABAR", "HMMA", "FFMA2",
       "MUFU", "SHFL", "SYNCS", "ELECT", "REDG", "ATOMG", "RED."]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip()[:110]
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in PAT:
            if re.search(r"\b" + re.escape(p), line):
                counts[cur][p] += 1
                total[p] += 1
    print("# " + os.path.relpath(LIB) + " : SASS mnemonic counts per kernel (sm_100a)")
    print("# total: " + ", ".join(f"{k} {v}" for k, v in total.most_common()))
    for fn, c in counts.items():
        hot = {k: v for k, v in c.items() if k in ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UCGABAR", "HMMA", "FFMA2")}
        if hot:
            print(f"{fn}\n    " + ", ".join(f"{k} {v}" for k, v in sorted(hot.items())))


if __name__ == "__main__":
    main()
