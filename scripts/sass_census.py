"""SASS census of libavc_b200.so: instruction count per kernel and its tensor-core / TMEM / TMA mnemonics.
    cuobjdump -sass autoformer_b200/libavc_b200.so | python scripts/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import re
import subprocess
import sys

KEYS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "UCGABAR", "STAS", "UTMAPF")
cur = None
cnt = collections.defaultdict(collections.Counter)
tot = collections.Counter()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        tot[cur] += 1
        for key in KEYS:
            if op.startswith(key):
                cnt[cur][key + (".2CTA" if ".2CTA" in op else "")] += 1
print("# cuobjdump -sass autoformer_b200/libavc_b200.so (final build of round 2): SASS instruction count per kernel and the tensor-core /")
print("# TMEM / TMA mnemonics in it (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor load /")
print("# store, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, STAS = st.async to a cluster peer, UCGABAR = cluster barrier)")
names = subprocess.run(["c++filt"], input="\n".join(tot), capture_output=True, text=True).stdout.splitlines()
for k, name in sorted(zip(tot, names), key=lambda kn: -tot[kn[0]]):
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("avc::", "")
    print(f"{tot[k]:6d}  {name:48s} " + ", ".join(f"{a} {b}" for a, b in sorted(cnt[k].items())))
