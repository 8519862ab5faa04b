"""Opcode evidence per kernel of libmmcodec.so (run in the authoring container, no GPU needed):
    python profiles/sass_summary.py > profiles/r02_sass_summary.txt
Counts the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTMALDG = TMA tensor loads, UBLKCP = bulk copies, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path."""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = glob.glob(os.path.join(ROOT, "165-*", "mmcodec", "libmmcodec.so"))[0]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "MUFU", "LDG", "STG", "LDS", "STS", "BAR"]
cur, per = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["total"] += 1
        base = op.split(".")[0]
        if base in KEYS:
            per[cur][base] += 1
        if op.startswith("UTCHMMA.2CTA"):
            per[cur]["UTCHMMA.2CTA"] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(lib, ROOT)}: {len(per)} kernels, sm_100a SASS (cuobjdump -sass)")
tot = collections.Counter()
for (name, c), dn in zip(per.items(), demangle):
    tot.update(c)
    short = re.sub(r"\(.*", "", dn).replace("void mmc::", "").replace("mmc::", "")
    cols = " ".join(f"{k}={c[k]}" for k in KEYS if c[k])
    print(f"{short:70s} instr={c['total']:6d} {cols}")
print("# library totals: " + " ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
