"""Prints ms/step and the per-phase device times of one bench.py JSON line read from stdin (A/B helper for gpurun logs)."""
import json
import sys

line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
ph = d.get("phases_ms", {})
print("ms/step %.3f  Mtok/s %.2f  e2e %.2f  | " % (d["ms_per_step"], d["value"] / 1e6, d.get("e2e", {}).get("value", 0) / 1e6)
      + " ".join("%s=%.2f" % (k.replace("proj_", "").replace("recurrent", "rec")[:12], v) for k, v in ph.items()))
