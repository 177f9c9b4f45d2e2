#!/usr/bin/env python
"""One-line digest of a bench.py JSON line on stdin (tuning sessions): tools/bench_brief.py <label>"""
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
e = d.get("e2e", {})
print(" ".join(sys.argv[1:]), "| value", round(d["value"]), "ms", round(d["ms_per_step"], 2), "| e2e", round(e.get("value", 0)),
      "ms", round(e.get("ms_per_step", 0), 2), "| identical", e.get("bytes_identical_to_device_resident_run"),
      "| kernel ms", round(d.get("roofline", {}).get("launch_ms", 0), 2))
