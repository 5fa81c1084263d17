#!/usr/bin/env python
"""Opcode histogram of the SASS of one object file / shared library, per kernel: the evidence that a kernel is a
tcgen05 / TMEM / TMA kernel (UTC*MMA, LDTM / STTM, UTMALDG / UBLKCP, UTCBAR) or a DMMA one -- B200_PROFILING.md
"What proves a Blackwell-native kernel".  Runs here (no GPU needed).

    python tools/sass_hist.py build/obj/gemm_f32_tc.o [kernel-regex] > profiles/sass_gemm_f32_tc_r02.txt
"""
import collections
import re
import subprocess
import sys

KEY = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCCP",
       "DMMA", "HMMA", "IMMA", "LDGSTS", "SYNCS", "LDS", "STS", "LDG", "STG", "IMAD", "FFMA", "DFMA", "MUFU", "BAR", "RED", "ATOM")


def main():
    obj = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {obj}: opcode counts per kernel (static instructions)")
    for (name, cnt), dn in zip(kernels.items(), demangle):
        if pat and not pat.search(dn):
            continue
        total = sum(cnt.values())
        print(f"\n== {dn}\n   {total} instructions")
        shown = [(k, cnt[k]) for k in KEY if cnt.get(k)]
        print("   " + "  ".join(f"{k} {v}" for k, v in shown))
        rest = [(k, v) for k, v in cnt.most_common(12) if k not in KEY]
        print("   other (top): " + "  ".join(f"{k} {v}" for k, v in rest))


if __name__ == "__main__":
    main()
