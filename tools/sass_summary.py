#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libbci.so, how many of the instructions that prove a
Blackwell-native path it contains (B200_PROFILING.md "What proves a Blackwell-native kernel").
    python tools/sass_summary.py > profiles/r2_sass_summary.md"""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "bee_code_interpreter_b200" / "libbci.so"
MARKS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "LDG.256", "STG.256", "LDG.128", "STG.128",
         "LDGSTS", "HMMA", "IMAD.WIDE", "SYNCS", "ATOMS", "SHFL", "MATCH"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    demangle = {}
    per = defaultdict(Counter)
    total = Counter()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["_n"] += 1
            parts = op.split(".")
            for mark in MARKS:
                mp = mark.split(".")
                # "LDG.256" = an LDG whose modifiers include 256 (SASS spells it LDG.E.NA.ENL2.256)
                if parts[0] == mp[0] and all(x in parts[1:] for x in mp[1:]):
                    per[cur][mark] += 1
                    total[mark] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    def short(n: str) -> str:
        n = n.replace("(anonymous namespace)::", "").replace("bci::", "")
        if n.endswith(")"):                      # drop the parameter list: the "(" that matches the final ")"
            depth = 0
            for i in range(len(n) - 1, -1, -1):
                depth += n[i] == ")"
                depth -= n[i] == "("
                if depth == 0:
                    n = n[:i]
                    break
        n = re.sub(r"\((?:int|bool|unsigned int)\)", "", n)       # <(int)1, (bool)0> -> <1, 0>
        return n[5:] if n.startswith("void ") else n

    for k, n in zip(per, names):
        demangle[k] = short(n)
    arch = subprocess.run(["cuobjdump", "-lelf", str(LIB)], capture_output=True, text=True).stdout
    print("# SASS summary of libbci.so (round 2)\n")
    print("`cuobjdump -lelf`: " + ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))) + " only.\n")
    print("Totals over all kernels: " + ", ".join(f"{k} {v}" for k, v in total.most_common()) + "\n")
    print("| kernel | instructions | " + " | ".join(MARKS) + " |")
    print("|---|---:|" + "---:|" * len(MARKS))
    for k in sorted(per, key=lambda x: demangle[x]):
        c = per[k]
        if not any(c[m] for m in MARKS):
            continue
        print(f"| `{demangle[k][:90]}` | {c['_n']} | " + " | ".join(str(c[m]) if c[m] else "" for m in MARKS) + " |")


if __name__ == "__main__":
    main()
