"""Per-kernel SASS evidence for profiles/: counts of the mnemonics that prove the Blackwell-native path (tcgen05 -> UTC*MMA, TMEM
loads -> LDTM, TMA -> UTMALDG / UBLKCP, ...) and of the legacy tensor path (HMMA) in every kernel of libvcla.so.

    python tools/sass_opcodes.py > profiles/sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "visual-chinese-llama-alpaca_b200", "lib", "libvcla.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "UBLKCP", "UTCATOMSWS", "SYNCS", "HMMA", "LDGSTS", "LDSM", "MUFU", "UCGABAR", "REDUX", "ATOMS", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts, order, cur, i = {}, [], None, 0
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*$", "", names[i].replace("(anonymous namespace)::", "")).replace("vcla::", "")
            i += 1
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if cur and m:
            counts[cur][m.group(1)] += 1
            counts[cur]["_all"] += 1
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a): instruction counts per kernel; UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld,")
    print("# UTMALDG / UTMAPF = TMA tensor load / prefetch, UBLKCP = cp.async.bulk, SYNCS = mbarrier, UCGABAR = cluster barrier, HMMA = legacy mma.sync")
    print(f"{'kernel':64s} {'instr':>6s} " + " ".join(f"{o:>8s}" for o in OPS))
    for k in order:
        c = counts[k]
        print(f"{k[:64]:64s} {c['_all']:6d} " + " ".join(f"{sum(v for n, v in c.items() if n.startswith(o)):8d}" for o in OPS))


if __name__ == "__main__":
    main()
