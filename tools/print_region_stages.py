"""Print (variant, stage times) of tools/bench_regions.py JSON lines read from stdin (development helper for option sweeps)."""
import json
import sys

for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line)
        print(d.get("variant", ""), {k: round(v, 4) for k, v in d["ms"].items()})
