"""Counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use, per kernel of libcloudaae_b200.so
(cuobjdump -sass; runs without a GPU).   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cloudaae_b200", "lib", "libcloudaae_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
MN = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "REDUX", "ATOMS.CAST", "HMMA", "FFMA", "DFMA", "ACQBULK", "UBLKCP"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", name).replace("void ", "").replace("caae::", "")
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    mm = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if mm:
        op = mm.group(1)
        per[cur]["_total"] += 1
        for k in MN:
            if op.startswith(k):
                per[cur][k] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(lib, ROOT)} (sm_100a; cuobjdump -sass)")
print(f"# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG/UTMASTG/UTMAREDG = TMA tensor load/store/reduce,")
print(f"# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, REDUX = warp reduce, ATOMS.CAST = shared-memory CAS (float atomics)")
hdr = ["kernel", "instr"] + MN
print(" | ".join(hdr))
tot = collections.Counter()
for k, c in per.items():
    if not any(c[m] for m in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG", "REDUX", "ATOMS.CAST", "DFMA")) and "--all" not in sys.argv:
        continue
    print(" | ".join([k[:70], str(c["_total"])] + [str(c[m]) for m in MN]))
    tot.update(c)
print(" | ".join(["TOTAL (listed kernels)", str(tot["_total"])] + [str(tot[m]) for m in MN]))
