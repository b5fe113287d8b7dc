#!/usr/bin/env python
"""Static resource usage of libgracing.so (no GPU): `cuobjdump -res-usage` per kernel + SASS mnemonic counts -> markdown on stdout.

  python tools/static_resource_usage.py "final round-2 build" > profiles/r2_static_resource_usage.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "generalizableracing_b200", "libgracing.so")


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "current build"
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    kernels = []
    name = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            kernels.append((name,) + tuple(int(x) for x in m.groups()))
            name = None
    dem = subprocess.run(["c++filt"], input="\n".join(k[0] for k in kernels), capture_output=True, text=True).stdout.splitlines()
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    ops = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops[m.group(1)] += 1
    pick = lambda pre: sum(v for k, v in ops.items() if k.startswith(pre))
    print(f"# Static resource usage of `libgracing.so` (sm_100a), {what}\n")
    print("`python tools/static_resource_usage.py` = `cuobjdump -res-usage` + `cuobjdump -sass` of the library `__graft_entry__.build()` produces (nvcc 12.9, "
          "`-gencode arch=compute_100a,code=sm_100a -lineinfo`); no GPU needed.\n")
    print(f"{len(kernels)} kernels; registers per thread: max {max(k[1] for k in kernels)}; kernels with a stack frame: {sum(1 for k in kernels if k[2] > 0)}; "
          f"kernels with LOCAL > 0: {sum(1 for k in kernels if k[4] > 0)}.\n")
    fams = [("UTCHMMA", "tcgen05.mma"), ("LDTM", "tcgen05.ld"), ("UTCBAR", "tcgen05.commit"), ("SYNCS", "mbarrier ops"), ("ELECT", "elect.sync"), ("LDGSTS", "cp.async"),
            ("LDG.E.128", "16-byte global loads"), ("STG.E.128", "16-byte global stores"), ("LDL", "local (spill) loads"), ("STL", "local (spill) stores"),
            ("ACQBULK", "griddepcontrol.wait"), ("REDG", "global reductions"), ("STG.E.STRONG.SYS", "st.release.sys (peer flags)"), ("LDG.E.STRONG.SYS", "ld.acquire.sys (peer flags)"), ("LDG.E.128.STRONG.SYS", "volatile 16-byte peer loads")]
    print("SASS mnemonic counts over the whole library: " + ", ".join(f"`{p}` {pick(p)} ({d})" for p, d in fams) + ".\n")
    print("| kernel | REG | STACK | static SHARED | LOCAL |\n|---|---:|---:|---:|---:|")
    for (n, reg, stack, shared, local), d in sorted(zip(kernels, dem), key=lambda kd: (-kd[0][1], kd[1])):
        d = re.sub(r"\(.*", "", d)
        print(f"| `{d}` | {reg} | {stack} | {shared} | {local} |")


if __name__ == "__main__":
    main()
