"""SASS census of libspai_b200.so: per kernel, how many instructions of the mnemonics that identify the
Blackwell / Hopper+ mechanisms this library relies on (profiles/sass_census.txt).

    python tools/sass_census.py > profiles/sass_census.txt

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st (tensor memory), UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async, ATOMS / RED = shared / global atomics."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gflownet_spai_b200", "libspai_b200.so")
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS", "LDGSTS", "ATOMS", "ATOMG", "RED", "MUFU", "FFMA", "DFMA",
        "DADD", "SHFL", "VOTE", "LDS", "STS", "LDG", "STG", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kern = None
    counts = collections.OrderedDict()
    total = collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("spai::", "")
            kern = name
            counts.setdefault(kern, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and kern:
            op = m.group(1)
            counts[kern]["_all"] += 1
            for k in KEYS:
                if op.startswith(k):
                    counts[kern][k] += 1
                    total[k] += 1
                    break
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels, {sum(c['_all'] for c in counts.values())} SASS instructions")
    print("# whole library:", " ".join(f"{k}={total[k]}" for k in KEYS if total[k]))
    print(f"{'kernel':70s} {'instrs':>7s}  " + " ".join(f"{k:>7s}" for k in KEYS[:14]))
    for name, c in counts.items():
        print(f"{name[:70]:70s} {c['_all']:7d}  " + " ".join(f"{c[k]:7d}" for k in KEYS[:14]))


if __name__ == "__main__":
    main()
