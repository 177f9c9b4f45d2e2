#!/usr/bin/env python
"""Where are the local-memory (spill) instructions of k_pcg_cluster<CL> relative to its iteration loop?
The loop is found as the backward branch that spans the halo-barrier wait (SYNCS.PHASECHK on mbarP);
prints, per cluster size, the LDL/STL count inside and outside that range.

    python tools/sass_spills.py [build/k_pcg_cluster.o]
"""
import re, subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else "build/k_pcg_cluster.o"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)[1:]
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    if "k_pcg_cluster" not in name:
        continue
    ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", f)]
    waits = [a for a, t in ins if "SYNCS.PHASECHK" in t]
    best = None
    for a, t in ins:
        m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a:
                n = sum(1 for w in waits if tgt <= w <= a)
                if n >= 6 and (best is None or (a - tgt) < (best[1] - best[0])):   # smallest loop with the SpMV's waits
                    best = (tgt, a)
    spills = [(a, t.split()[0]) for a, t in ins if re.match(r"(@!?P\d\s+)?(LDL|STL)", t)]
    inside = [s for s in spills if best and best[0] <= s[0] <= best[1]]
    print("%s: loop 0x%x..0x%x, %d instructions; spill instructions inside %d, outside %d %s"
          % (re.search(r"ILi(\d)E", name).group(0), best[0], best[1], sum(1 for a, _ in ins if best[0] <= a <= best[1]),
             len(inside), len(spills) - len(inside), [hex(a) for a, _ in inside]))
