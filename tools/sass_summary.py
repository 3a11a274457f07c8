"""Per-kernel SASS evidence of libelo_b200.so (no GPU needed): counts of the Blackwell-specific instructions.

    python tools/sass_summary.py > profiles/r2/sass_r2.txt

UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA bulk
copy), UTMALDG / UTMASTG = cp.async.bulk.tensor (none: the tiles this path moves are gathers and 12-byte-stride rows,
which a tensor map cannot describe; weights and the index kernel's xyz tile ride the un-tiled UBLKCP -- DESIGN.md
section 4.1), SYNCS = mbarrier ops, REDUX / VOTE / SHFL = warp
collectives of the searches."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "efficientlo-net_b200", "csrc", "libelo_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "REDUX", "VOTE",
         "SHFL", "LDS", "STS", "LDG", "STG", "ATOM", "RED", "BAR", "ACQBULK", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    kernels[cur][w] += 1
                    break
    demangle = subprocess.run(["c++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print("libelo_b200.so, sm_100a: SASS instruction counts per kernel (static, not executed counts)\n")
    cols = [w for w in WATCH if any(k[w] for k in kernels.values())]
    print("%-96s %7s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
    tot = collections.Counter()
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*\)$", "", dm).replace("elo::", "").replace("void ", "")
        print("%-96s %7d " % (short[:96], c["_total"]) + " ".join("%7d" % c[w] for w in cols))
        tot.update(c)
    print("%-96s %7d " % ("TOTAL", tot["_total"]) + " ".join("%7d" % tot[w] for w in cols))


if __name__ == "__main__":
    main()
