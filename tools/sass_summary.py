#!/usr/bin/env python3
"""Per-kernel SASS evidence of the in-tree library: counts of the instructions that prove the hardware paths
(DMMA = FP64 tensor-core MMA, UBLKCP = TMA bulk copy, SYNCS = mbarrier, LDGSTS = cp.async, DFMA/DMUL/DADD = FP64 pipe,
LDS/STS shared memory, BAR barriers) plus registers / spills from the ptxas logs.

  python tools/sass_summary.py > profiles/sass_r02.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "polydeal_b200", "lib", "libpolydeal_b200.so")
OPS = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "ATOM", "RED"]


def demangle(names):
    # internal-linkage kernels are emitted as __nv_static_<n>__<hash>_<file>__<mangled name>
    inner = [re.search(r"(_ZN\w+)", n).group(1) if re.search(r"(_ZN\w+)", n) else n for n in names]
    out = subprocess.run(["c++filt"], input="\n".join(inner), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            for key in OPS:
                if op == key or op.startswith(key + "."):
                    counts[cur][key] += 1
            counts[cur]["total"] += 1
    regs = {}
    for log in sorted(glob.glob(os.path.join(ROOT, "polydeal_b200", "build", "*.ptxas.log"))):
        cur = None
        for line in open(log):
            m = re.search(r"Compiling entry function '([^']+)'", line)
            if m:
                cur = m.group(1)
                regs[cur] = [None, 0, 0]
            m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and cur:
                regs[cur][1], regs[cur][2] = int(m.group(1)), int(m.group(2))
            m = re.search(r"Used (\d+) registers", line)
            if m and cur:
                regs[cur][0] = int(m.group(1))
    names = demangle(list(counts))
    short = lambda n: re.sub(r"\(.*", "", re.sub(r"pd::\(anonymous namespace\)::", "", names.get(n, n)))[:64]
    print(f"# SASS summary of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); regs / spill bytes from ptxas -v")
    print(f"{'kernel':64s} {'instr':>7s} " + " ".join(f"{k:>6s}" for k in OPS) + "   regs spill_st spill_ld")
    tot = collections.Counter()
    for n, c in counts.items():
        r = regs.get(n, ("", "", ""))
        print(f"{short(n):64s} {c['total']:7d} " + " ".join(f"{c[k]:6d}" for k in OPS) + f"   {r[0]!s:>4s} {r[1]!s:>8s} {r[2]!s:>8s}")
        tot.update(c)
    print(f"{'TOTAL':64s} {tot['total']:7d} " + " ".join(f"{tot[k]:6d}" for k in OPS))
    print("# tcgen05 (UTCMMA / LDTM) is absent by design: tcgen05 has no f64 kind; the FP64 tensor path on sm_100a is mma.sync -> DMMA.8x8x4")
    for pat in ("UTC", "LDTM", "HMMA", "QGMMA"):
        print(f"# occurrences of {pat}: {len(re.findall(pat, sass))}")


if __name__ == "__main__":
    main()
