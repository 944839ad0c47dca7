"""Count the Blackwell-specific SASS mnemonics per kernel of libseedb200.so (cuobjdump -sass), the evidence
B200_PROFILING.md asks for: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor loads,
UTCBAR = tcgen05.commit, HMMA = legacy mma.sync, LDGSTS = cp.async, ACQBULK = griddepcontrol.wait (PDL).
usage: python tools/sass_evidence.py > profiles/r01_sass_evidence.md"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "seed_b200", "libseedb200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UBLKCP|UTCBAR|UTCATOMSWS|LDTM|STTM|HMMA|LDGSTS|SYNCS|ACQBULK|MUFU\.EX2)\b")
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    if cur:
        for t in pat.findall(line):
            counts[cur][t] += 1
def demangle(n):
    out = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    return re.sub(r"\(.*$", "", out).replace("void ", "").replace("sb::", "")
cols = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "LDGSTS", "ACQBULK", "MUFU.EX2"]
print("# SASS evidence (cuobjdump -sass seed_b200/libseedb200.so, sm_100a): static instruction counts per kernel\n")
print("UTCHMMA = `tcgen05.mma` kind::f16, UTCBAR = `tcgen05.commit`, UTCATOMSWS = TMEM alloc/dealloc, LDTM/STTM = `tcgen05.ld/st`,")
print("UTMALDG = `cp.async.bulk.tensor` (TMA), SYNCS = mbarrier ops, HMMA = legacy `mma.sync`, LDGSTS = `cp.async`,")
print("ACQBULK = `griddepcontrol.wait` (programmatic dependent launch).  No HGMMA/wgmma anywhere (sm_90-only).\n")
print("| kernel | " + " | ".join(cols) + " |")
print("|---|" + "---:|" * len(cols))
seen = set()
for k, c in counts.items():
    name = demangle(k)
    if name in seen:
        continue
    seen.add(name)
    if not any(c[x] for x in cols):
        continue
    print(f"| `{name}` | " + " | ".join(str(c[x]) if c[x] else "" for x in cols) + " |")
