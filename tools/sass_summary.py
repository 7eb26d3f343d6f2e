#!/usr/bin/env python
"""profiles/rNN_sass_summary.txt: per-kernel counts of the Blackwell-specific SASS instructions in libscan_b200.so.

    cuobjdump -sass scan_b200/libscan_b200.so > /tmp/sass.txt && python tools/sass_summary.py /tmp/sass.txt > profiles/r02_sass_summary.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit,
SYNCS = mbarrier operations, REDG = red.global (the accumulator drains).  Runs in the build container (no GPU needed)."""
import re
import subprocess
import sys


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


def main(path):
    txt = open(path).read()
    rows = []
    for p in re.split(r"\n\s*Function : ", txt)[1:]:
        name = p.split("\n", 1)[0].strip()
        cnt = lambda pat: len(re.findall(pat, p))  # noqa: E731
        rows.append((name, dict(mma=cnt(r"\bUTC[A-Z]*MMA"), ldtm=cnt(r"\bLDTM"), sttm=cnt(r"\bSTTM"), tma=cnt(r"\bUTMALDG"),
                                bar=cnt(r"\bUTCBAR"), syncs=cnt(r"\bSYNCS"), red=cnt(r"\bREDG\."), total=cnt(r"\n\s+/\*[0-9a-f]{4}\*/"))))
    print("# SASS summary of scan_b200/libscan_b200.so (sm_100a): tensor-core / tensor-memory / TMA instructions per kernel")
    print("# UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load, UTCBAR = tcgen05.commit, SYNCS = mbarrier, REDG = red.global")
    print("%-92s %7s %5s %5s %8s %7s %6s %5s %7s" % ("kernel", "UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTCBAR", "SYNCS", "REDG", "instrs"))
    for n, d in sorted([r for r in rows if r[1]["mma"] or r[1]["tma"]], key=lambda r: -r[1]["mma"]):
        print("%-92s %7d %5d %5d %8d %7d %6d %5d %7d" % (re.sub(r"\(.*", "", demangle(n))[:92], d["mma"], d["ldtm"], d["sttm"], d["tma"],
                                                        d["bar"], d["syncs"], d["red"], d["total"]))
    print()
    print("# all %d kernels in the library (name, SASS instruction count)" % len(rows))
    for n, d in sorted(rows, key=lambda r: r[0]):
        print("%-100s %7d" % (re.sub(r"\(.*", "", demangle(n))[:100], d["total"]))


if __name__ == "__main__":
    main(sys.argv[1])
