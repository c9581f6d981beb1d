"""Development helper: registers / spills per kernel from an `nvcc -Xptxas -v` log.  python tools/regs.py log [filter]"""
import re, sys
cur, spill = None, ""
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for line in open(sys.argv[1]):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
    if "spill" in line:
        spill = re.sub(r"\s+", " ", line.strip())
    m2 = re.search(r"Used (\d+) registers", line)
    if m2 and cur and flt in cur:
        t = re.findall(r"IL[ib](\d+)E|L[ib](\d+)E", cur)
        print(re.sub(r"^_ZN3fdb\d+", "", cur)[:28], [a or b for a, b in t], "regs", m2.group(1), "|", spill.replace("bytes", "B")[:60])
