#!/usr/bin/env python
"""SASS summary of libvfuse.so for profiles/: per kernel the instruction count and the mnemonics that identify the Blackwell paths
(tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, tcgen05.cp = UTCCP, TMA = UTMALDG / UTMASTG, mbarrier = SYNCS, spills = STL/LDL).

    python tools/sass_summary.py r02_v4        # writes profiles/r02_v4_sass_summary.txt (needs only cuobjdump, no GPU)
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "llm_quest_b200" / "libvfuse.so"
PAT = [("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM/STTM (tcgen05.ld/st)", r"\b(LDTM|STTM)"), ("UTCCP (tcgen05.cp)", r"\bUTCCP"),
       ("UTMALDG (TMA load)", r"\bUTMALDG"), ("UTMASTG (TMA store)", r"\bUTMASTG"), ("SYNCS (mbarrier)", r"\bSYNCS"),
       ("MUFU.EX2", r"\bMUFU\.EX2"), ("STL/LDL (spills)", r"\b(STL|LDL)\b"), ("multimem / STG.E.128", r"\bSTG\.E\.128")]

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
archs = sorted(set(re.findall(r"sm_\d+a?", subprocess.run(["cuobjdump", "--list-elf", str(LIB)], capture_output=True, text=True).stdout)))
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
out = [f"SASS summary of llm_quest_b200/libvfuse.so (cuobjdump -sass), cubin architectures: {archs}",
       "per kernel: SASS instruction count and the mnemonics that identify the Blackwell paths (B200_PROFILING.md table)", ""]
name, counts, n = None, None, 0
def flush():
    if name:
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem)
        out.append(f"{dem[:110]:112s} {n:5d} instr  " + ", ".join(f"{k} {v}" for k, v in counts.items() if v))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name, counts, n = m.group(1), collections.OrderedDict((k, 0) for k, _ in PAT), 0
        continue
    if name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        n += 1
        for k, p in PAT:
            if re.search(p, line):
                counts[k] += 1
flush()
(ROOT / "profiles" / f"{tag}_sass_summary.txt").write_text("\n".join(out) + "\n")
print("wrote", ROOT / "profiles" / f"{tag}_sass_summary.txt", len(out) - 3, "kernels")
