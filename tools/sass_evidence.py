#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell data path (cuobjdump -sass of the built library; needs no GPU):
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA tensor load),
UBLKCP = cp.async.bulk, SYNCS = mbarrier ops, REDG = red.global (x2 / x4 = vector REDs).
    python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "consistent_depth_b200", "lib", "libcvd_sm100.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
demangled = dict(zip(re.findall(r"Function : (\S+)", sass), names))
counts, fn = collections.defaultdict(collections.Counter), None
pat = re.compile(r"\b(UTCHMMA|LDTM|UTCBAR|UTMALDG|UBLKCP|SYNCS|REDG|RED)\b[.\w]*")
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = pat.search(line)
    if m and fn:
        op = m.group(0)
        key = op.split(".")[0]
        if key == "UTMALDG":
            key = op                                    # keep the dimensionality (UTMALDG.5D)
        if key in ("REDG", "RED"):
            key = "REDG" + (".x4" if "128" in op or "F32x4" in op else ".x2" if ".64" in op or "F32x2" in op else "")
        counts[fn][key] += 1
print(f"# {os.path.relpath(lib, ROOT)}: SASS mnemonic counts per kernel (static instruction counts)")
for fn in sorted(counts, key=lambda f: demangled.get(f, f)):
    name = re.sub(r"\(anonymous namespace\)::", "", demangled.get(fn, fn))
    name = re.sub(r"\(.*", "", name)
    c = counts[fn]
    if not any(k.startswith(("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "REDG")) for k in c):
        continue
    print(f"{name:60s} " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())))
