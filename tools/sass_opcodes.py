"""cuobjdump -sass of the built library -> opcode counts per kernel (the tcgen05 / TMA / mbarrier mnemonics of B200_PROFILING.md).
usage: cuobjdump -sass dex-tts_b200/dexb200/libdexb200.so | python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

WANT = ["UTCHMMA", "UTMALDG", "UTMASTG", "UTMACCTL", "LDTM", "STTM", "UTCBAR", "SYNCS", "MEMBAR", "ATOMG", "REDG"]
kern, counts, instrs = None, collections.OrderedDict(), {}
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        instrs[kern] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        instrs[kern] += 1
        op = m.group(1)
        if op in WANT:
            counts[kern][op] += 1
names = list(counts)
try:
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
except Exception:
    dem = names
print("# r02 -- SASS opcode counts per kernel of dex-tts_b200/dexb200/libdexb200.so (cuobjdump -sass, sm_100a), final build")
print("# tcgen05.mma = UTCHMMA, TMA load = UTMALDG, TMA store = UTMASTG, tcgen05.ld/st = LDTM/STTM, tcgen05.commit = UTCBAR, mbarrier = SYNCS")
print(f"{'kernel':110s} {'instrs':>6s}  opcodes")
tot = collections.Counter()
for k, d in zip(names, dem):
    c = counts[k]
    if not c:
        continue
    tot.update(c)
    d = re.sub(r"\(.*", "", d)
    print(f"{d[:110]:110s} {instrs[k]:6d}  " + " ".join(f"{o}={c[o]}" for o in WANT if c[o]))
print("\nTOTAL " + " ".join(f"{o}={tot[o]}" for o in WANT if tot[o]))
