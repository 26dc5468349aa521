"""Compact SASS evidence for the shipped library (runs on the CPU box: cuobjdump only): per kernel the register /
stack / shared-memory budget and the counts of the mnemonics that show which Blackwell mechanisms the code uses
(TMA tensor loads, mbarrier transactions, programmatic dependent launch, packed f32x2 arithmetic, warp reductions).
    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rpg_open_remode_b200 import _build

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
CUFILT = shutil.which("cu++filt") or "/usr/local/cuda/bin/cu++filt"
MNEMONICS = ["UTMALDG", "SYNCS", "ACQBULK", "PREEXIT", "FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "LDS", "STS",
             "LDG", "STG", "ATOMS", "ATOMG", "REDUX", "SHFL", "VOTE", "BAR", "MEMBAR", "ERRBAR", "LDL", "STL"]


def main():
    lib = _build.build_cuda()
    sass = subprocess.run([CUOBJDUMP, "-sass", lib], capture_output=True, text=True, check=True).stdout
    res = subprocess.run([CUOBJDUMP, "-res-usage", lib], capture_output=True, text=True, check=True).stdout
    usage = dict(re.findall(r"Function (\S+):\s*\n\s*(.*)", res))
    print("# r02 -- SASS summary of rpg_open_remode_b200/librmd_b200.so (sm_100a), written by tools/sass_summary.py\n")
    print("Counts are static instruction counts per kernel.  `UTMALDG` = `cp.async.bulk.tensor` (TMA), `SYNCS` = mbarrier "
          "arrive / try_wait, `ACQBULK` = `griddepcontrol.wait` (programmatic dependent launch), `FFMA2/FMUL2/FADD2` = packed "
          "`f32x2` arithmetic, `LDL/STL` = local memory (spills).\n")
    print("| kernel | registers | stack B | instructions | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|---|---|" + "---|" * len(MNEMONICS))
    for part in sass.split("Function : ")[1:]:
        name = part.split("\n", 1)[0].strip()
        ops = re.findall(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", part, re.M)
        if len(ops) < 40:
            continue
        pretty = subprocess.run([CUFILT, name], capture_output=True, text=True).stdout.strip() or name
        pretty = re.sub(r"\(.*$", "", pretty.replace("(int)", "")).replace("void ", "").replace("rmdb::", "").replace("(anonymous namespace)::", "")
        u = usage.get(name, "")
        reg = re.search(r"REG:(\d+)", u)
        stack = re.search(r"STACK:(\d+)", u)
        counts = [sum(1 for o in ops if o == m or (m in ("LDG", "STG", "LDS", "STS", "LDL", "STL") and o.startswith(m))) for m in MNEMONICS]
        print(f"| `{pretty}` | {reg.group(1) if reg else '?'} | {stack.group(1) if stack else '?'} | {len(ops)} | " +
              " | ".join(str(c) for c in counts) + " |")


if __name__ == "__main__":
    main()
