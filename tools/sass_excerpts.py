#!/usr/bin/env python
"""SASS excerpts of the dominant kernels (cuobjdump -sass of csrc/gicp.o and csrc/fgr.o) -> profiles/<name>.md:
opcode histogram, the tensor-core / packed-FP32 / async-copy mnemonics that prove the design, and the inner loops.
usage: python tools/sass_excerpts.py profiles/r2_01_sass_excerpts.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "multi-scale-pointcloud-registration_b200", "csrc")


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj)], capture_output=True, text=True).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            funcs[name].append(line.rstrip())
    return funcs


def opc(line):
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    return m.group(2) if m else None


def count(ins, prefixes):
    return ", ".join(f"{x} x{sum(1 for l in ins if opc(l) and opc(l).startswith(x))}" for x in prefixes)


def around(ins, prefix, before, after):
    idx = [i for i, l in enumerate(ins) if opc(l) and opc(l).startswith(prefix)]
    if not idx:
        return "(none)"
    return "\n".join(ins[max(0, idx[0] - before):min(len(ins), idx[0] + after)])


def main():
    out = ["# SASS excerpts of the dominant kernels (cuobjdump -sass, sm_100a)\n",
           "Written by `python tools/sass_excerpts.py <this file>` from `csrc/gicp.o` and `csrc/fgr.o` "
           "(built by `make -C multi-scale-pointcloud-registration_b200/csrc`).\n"]
    g = sass("gicp.o")
    ins = g[[n for n in g if "gicp_persistent_kernelILi1E" in n][0]]
    cnt = collections.Counter(opc(l).split(".")[0] for l in ins if opc(l))
    out.append(f"## gicp_persistent_kernel<1> ({len(ins)} instructions, {len(ins) * 16 // 1024} KB)\n")
    out.append("opcode histogram (top 24): " + ", ".join(f"{a} {b}" for a, b in cnt.most_common(24)) + "\n")
    out.append("tensor-core / packed-FP32 / async-copy evidence: " +
               count(ins, ("DMMA", "FFMA2", "FMNMX3", "LDGSTS", "LDS.128", "ATOMG", "NANOSLEEP", "SHFL")) + "\n")
    out.append("\n### the FP32 sweep (4-target steps of the chunk stream: 4 LDS.128 feed 12 FFMA2 and 4 FMNMX3)\n```\n" +
               around(ins, "FFMA2", 12, 60) + "\n```\n")
    out.append("\n### the float64 tensor-core Jacobian sums (mma.m8n8k4.f64: LDS.64 of a record entry and of its weight, "
               "DMUL, DMMA)\n```\n" + around(ins, "DMMA", 14, 26) + "\n```\n")
    out.append("\n### the work-queue wait (ticket, volatile poll of the ring slot, ATOMG.EXCH claim, done flag, back-off)\n```\n" +
               around(ins, "NANOSLEEP", 14, 8) + "\n```\n")
    f = sass("fgr.o")
    ins = f[[n for n in f if "feat_match_kernel" in n][0]]
    out.append(f"\n## feat_match_kernel<0> ({len(ins)} instructions)\n")
    out.append("tcgen05 / TMEM / TMA evidence: " +
               count(ins, ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "UTCATOMSWS", "FFMA", "FMNMX")) + "\n")
    out.append("\n### MMA issue (one elected thread: tcgen05.mma kind::tf32 into TMEM, commit to an mbarrier)\n```\n" +
               around(ins, "UTCHMMA", 10, 24) + "\n```\n")
    out.append("\n### epilogue (tcgen05.ld of the accumulator, FP32 fold)\n```\n" + around(ins, "LDTM", 4, 16) + "\n```\n")
    open(sys.argv[1], "w").write("\n".join(out))


if __name__ == "__main__":
    main()
