#!/usr/bin/env python
"""Where the samples of the cluster kernel's iteration loop fall, from the SASS source page of an ncu capture:
    ncu -i <rep> --page source --csv > src.csv ;  python profiles/hot_segments.py src.csv "<3>"
The loop is cut at its synchronisation instructions (BAR.SYNC, mbarrier waits/arrives, st.async, the reciprocal
of the division), each segment printed with its share of the warp-state samples, executed instructions and
shared-memory wavefronts."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
I = lambda x: int(x) if x not in ("", "-") else 0
for sec in secs:
    if want not in sec["name"].replace("(int)", ""):
        continue
    hdr, data = sec["hdr"], sec["data"]
    isrc, isamp, iex, iwf = (hdr.index(k) for k in ("Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared"))
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(I(r[isamp]) for r in data)
    print(sec["name"], "samples", tot, "warp instructions", sum(I(r[iex]) for r in data), "shared wavefronts", sum(I(r[iwf]) for r in data))
    big = sorted(I(r[iex]) for r in data if I(r[iex]) > 0)
    hot = [i for i, r in enumerate(data) if I(r[iex]) > 0.5 * big[len(big) // 2]]   # the iteration loop: at least half the median count
    seg, label = [], "loop top"

    def flush():
        if seg:
            s = sum(I(r[isamp]) for r in seg)
            top = sorted(((sum(I(r[i]) for r in seg), hdr[i][6:]) for i in stall), reverse=True)[:3]
            print("%-58s n=%4d samples %5.1f%% instr %6.0fM wf %5dM  %s" % (label[:58], len(seg), 100.0 * s / tot, sum(I(r[iex]) for r in seg) / 1e6,
                  sum(I(r[iwf]) for r in seg) / 1e6, " ".join("%s %.0f%%" % (n, 100.0 * v / max(s, 1)) for v, n in top)))
    for i in range(hot[0], hot[-1] + 1):
        src = data[i][isrc].strip()
        if any(m in src for m in ("BAR.SYNC", "SYNCS.PHASECHK", "SYNCS.ARRIVE", "STAS", "MUFU.RCP64H")):
            flush()
            seg, label = [], src + " @%d" % i
        seg.append(data[i])
    flush()
