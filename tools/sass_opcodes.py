"""Per-kernel SASS opcode counts of librvsr_b200.so (cuobjdump -sass): the evidence that the hot kernels really are
tcgen05 / TMEM / TMA code.  UTCHMMA = tcgen05.mma (kind::f16), .2CTA = cta_group::2; LDTM = tcgen05.ld; UTMALDG = TMA
tensor load; UBLKCP = cp.async.bulk; SYNCS = mbarrier ops; UTCBAR = tcgen05.commit.

    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt      (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "realvsr_b200", "librvsr_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "HFMA2", "FFMA", "LDG", "STG", "LDS", "STS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    sub = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    names = iter(sub)
    cur, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names)
            cur = cur.replace("(anonymous namespace)::", "").replace("rvsr::", "")
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            c = counts[cur]
            c["total"] += 1
            base = op.split(".")[0]
            c[base] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                c["UTCHMMA.2CTA"] += 1
    print("# cuobjdump -sass realvsr_b200/librvsr_b200.so -- opcode counts per kernel (static instruction counts)")
    print("%-64s %7s " % ("kernel", "total") + " ".join("%12s" % k for k in KEYS))
    for k, c in counts.items():
        if c["total"] == 0:
            continue
        print("%-64s %7d " % (k[:64], c["total"]) + " ".join("%12d" % c[x] for x in KEYS))


if __name__ == "__main__":
    sys.exit(main())
