#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = TMA bulk copy,
HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async, SYNCS = mbarrier ops.

    python tools/sass_summary.py > profiles/sass_summary.txt      (cuobjdump on whisper_b200/lib/libwhisper_b200.so)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "whisper_b200", "lib", "libwhisper_b200.so")
PATTERNS = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
            ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("HMMA", r"\bHMMA"),
            ("LDGSTS", r"\bLDGSTS"), ("LDSM", r"\bLDSM"), ("SYNCS", r"\bSYNCS"), ("MUFU", r"\bMUFU"), ("BAR", r"\bBAR\.")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            counts[cur]["instructions"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", line):
            counts[cur]["instructions"] += 1
            for name, pat in PATTERNS:
                if re.search(pat, line):
                    counts[cur][name] += 1
    names = list(counts)
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        demangle[n] = d
    cols = ["instructions"] + [n for n, _ in PATTERNS]
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print(f"# {'kernel':<100} " + " ".join(f"{c:>8}" for c in cols))
    for n in names:
        short = re.sub(r"\(.*", "", demangle.get(n, n))
        short = short.replace("void wb::", "").replace("__nv_bfloat16", "bf16").replace("__half", "f16")
        print(f"{short[:102]:<102} " + " ".join(f"{counts[n][c]:>8}" for c in cols))


if __name__ == "__main__":
    sys.exit(main())
