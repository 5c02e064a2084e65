"""Per-kernel count of the SASS opcodes that prove a Blackwell-native path (B200_PROFILING.md, "What proves a
Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP.
    python scripts/sass_summary.py dir_b200/libdirb200.so > profiles/sass_opcodes_r2.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "dir_b200/libdirb200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pats = ["UTCHMMA.2CTA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCATOMSWS", "HMMA",
        "SYNCS", "FFMA", "MUFU"]
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("dirb200::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        cur = re.sub(r"\(.*", "", cur)
        per.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    for p in pats:
        if op.startswith(p):
            per[cur][p] += 1
            break
print(f"# SASS opcode counts per kernel of {so} (cuobjdump -sass); UTCHMMA = tcgen05.mma (kind::f16 and kind::tf32),")
print("# LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, HMMA = legacy mma.sync")
print(f"{'kernel':70s} " + " ".join(f"{p:>12s}" for p in pats))
tot = collections.Counter()
for k, c in per.items():
    if not any(c[p] for p in pats[:10]):
        continue
    print(f"{k[:70]:70s} " + " ".join(f"{c[p]:12d}" for p in pats))
    tot.update(c)
print(f"{'TOTAL (kernels listed above)':70s} " + " ".join(f"{tot[p]:12d}" for p in pats))
simt = [k for k, c in per.items() if not any(c[p] for p in pats[:10])]
print(f"# {len(simt)} further kernels use no tensor-core/TMA opcode (elementwise, packing, CUDA-core joint-space and fp32 A/B paths)")
