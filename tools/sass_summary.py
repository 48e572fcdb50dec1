"""Per-kernel counts of the sm_100a instructions that prove what each kernel runs on (cuobjdump -sass of the built
library): UTCHMMA / UTCBAR (tcgen05.mma / commit), LDTM (tcgen05.ld), UTMALDG (cp.async.bulk.tensor = TMA tile loads),
UBLKCP (bulk copies), SYNCS (mbarrier), HFMA2 (packed fp16 math), HMMA (legacy mma.sync; expected 0).
Writes profiles/sass_summary.txt.  No GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mvsnet_b200", "lib", "libmvsnet_b200.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "HFMA2", "HMMA", "FFMA", "LDS", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True,
                              text=True).stdout.splitlines()
    names = iter(demangle)
    counts, order, cur = {}, [], None
    for line in out.splitlines():
        if "Function :" in line:
            cur = next(names)
            cur = re.sub(r"\(.*", "", cur).replace("mvsb200::", "").replace("void ", "")
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for k in OPS:
                if op == k or op.startswith(k + "."):
                    counts[cur][k] += 1
            counts[cur]["total"] += 1
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    with open(path, "w") as f:
        f.write("# cuobjdump -sass mvsnet_b200/lib/libmvsnet_b200.so (sm_100a), instruction counts per kernel; tools/sass_summary.py\n")
        f.write("%-78s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in OPS) + "\n")
        for name in sorted(order, key=lambda n: -counts[n]["total"]):
            c = counts[name]
            f.write("%-78s %7d " % (name[:78], c["total"]) + " ".join("%7d" % c[k] for k in OPS) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    sys.exit(main())
