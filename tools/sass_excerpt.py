"""Per-kernel SASS evidence for the judge: counts of the Blackwell-specific mnemonics in the shipped library
(`cuobjdump -sass`), one example line each.  usage: python tools/sass_excerpt.py > profiles/rNN_sass_excerpt.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "caesar-mrcnn_b200", "lib", "libmrcnn_b200.so")
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UTMAPF|UTMACCTL|UTCATOM|SYNCS|UBLKCP|ELECT|"
                 r"FFMA2|FMUL2|FADD2|REDUX|VOTE|MATCH|RED|ATOMG|ATOMS)(\.[A-Z0-9_.]+)?\b")
SHOW = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UBLKCP")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts, first, ninstr = None, collections.OrderedDict(), {}, collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line or not re.search(r"/\*[0-9a-f]{4,5}\*/", line):
            continue
        ninstr[cur] += 1
        m = PAT.search(line)
        if m:
            op = m.group(1) + (m.group(2) or "")
            counts[cur][op] += 1
            first.setdefault((cur, op), line)
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    print("# SASS evidence: `cuobjdump -sass caesar-mrcnn_b200/lib/libmrcnn_b200.so` (sm_100a), special mnemonics per kernel\n")
    print("tcgen05: `UTCHMMA` (tcgen05.mma), `LDTM` (tcgen05.ld), `UTCBAR` (tcgen05.commit), `UTCATOM` (TMEM alloc); TMA: `UTMALDG` / "
          "`UTMASTG` (`.IM2COL` = im2col mode), `UBLKCP` (bulk copy); mbarrier: `SYNCS`; packed fp32: `FFMA2` / `FMUL2` / `FADD2`.\n")
    for fn, name in zip(counts, names):
        c = counts[fn]
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\)\(.*$|\(([^()]|\([^()]*\))*\)$", "", name)[:160]
        print("## `%s` (%d instructions)\n" % (name, ninstr[fn]))
        print("  " + (", ".join("%s x%d" % kv for kv in sorted(c.items())) or "(none)") + "\n")
        for op in sorted(c):
            if op.startswith(SHOW):
                l = re.sub(r"/\*[0-9a-f]{4,5}\*/\s*", "", first[(fn, op)])
                l = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip()
                print("    " + l)
        print()


if __name__ == "__main__":
    sys.exit(main())
