"""Per-kernel counts of the Blackwell-only SASS mnemonics in the built library (tcgen05 MMA = UTCHMMA / UTCQMMA, TMA loads and
stores = UTMALDG / UTMASTG, TMEM loads = LDTM, bulk reductions = UTMAREDG), from `cuobjdump -sass`.

    python tools/sass_summary.py [path/to/lib.so] > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "medmoe_b200", "lib", "libmedmoe_b200.so")
MNEMONICS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "REDG", "HMMA", "FFMA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        base = op.split(".")[0]
        counts[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            counts[cur]["UTCHMMA.2CTA"] += 1
    names = demangle(list(counts))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels; columns = SASS instruction counts per kernel")
    print("# " + " ".join(f"{m:>12s}" for m in MNEMONICS) + "  kernel")
    tot = collections.Counter()
    rows = []
    for k, c in counts.items():
        tot.update(c)
        short = re.sub(r"\(.*", "", names.get(k, k))
        rows.append((c["UTCHMMA"], "  " + " ".join(f"{c[m]:12d}" for m in MNEMONICS) + "  " + short))
    for _, r in sorted(rows, key=lambda t: -t[0]):
        print(r)
    print("# total")
    print("  " + " ".join(f"{tot[m]:12d}" for m in MNEMONICS))
    tc = sum(1 for c in counts.values() if c["UTCHMMA"])
    print(f"# kernels issuing tcgen05.mma: {tc}; using TMA loads: {sum(1 for c in counts.values() if c['UTMALDG'])}; "
          f"legacy mma.sync (HMMA) kernels: {sum(1 for c in counts.values() if c['HMMA'])}")


if __name__ == "__main__":
    main()
