"""Per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics of libfd_b200.so (cuobjdump -sass).
python tools/sass_extract.py > profiles/r2_sass_tcgen05_tma_vN.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pytorch-face-detection-from-scratch_b200", "libfd_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "UTCCP")
per, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for k in KEYS:
        if op.startswith(k):
            name = k
            if ".2CTA" in op:
                name += ".2CTA" + (".MULTICAST" if ".MULTICAST" in op else "")
            cur[name] += 1
print("# cuobjdump -sass libfd_b200.so (sm_100a): per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics")
print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit (.2CTA.MULTICAST = multicast commit of a CTA pair),")
print("# UTCATOMSWS = tcgen05.alloc/dealloc, UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce (.2CTA = cta_group::2 load signalling the leader's barrier), SYNCS = mbarrier ops")
tot = collections.Counter()
for fn, c in per.items():
    if not any(k.startswith(("UTCHMMA", "UTMALDG", "LDTM")) for k in c):
        continue
    name = demangle(fn).replace("(anonymous namespace)::", "").replace("void ", "")
    name = name[:name.index("(")] if "(" in name else name
    print(f"{name:52s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
    tot.update(c)
print("total".ljust(52), " ".join(f"{k}={v}" for k, v in sorted(tot.items())))
