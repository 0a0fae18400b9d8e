#!/usr/bin/env python
"""Print the interesting fields of a bench.py JSON line (file argument or stdin)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
lines = [x for x in txt.splitlines() if x.startswith("{")]
if not lines:
    print("no JSON line found; tail of the log:\n" + txt[-2000:])
    sys.exit(1)
d = json.loads(lines[-1])
print(f"{d['metric']}: {d['value']:.4g} {d['unit']}  ({d['ms_per_step']:.3f} ms/step, n_gpus {d['n_gpus']})")
if d.get("e2e"):
    print("  e2e:", {k: (f"{v:.4g}" if isinstance(v, float) else v) for k, v in d["e2e"].items() if k != "api"})
r = d.get("roofline")
if r:
    print(f"  roofline: {r['achieved']:.1f} {r['unit']} = {r['frac']:.3f} of {r['peak']}, {r['ms_per_launch']:.3f} ms/launch, share {r['share_of_step']:.3f}")
    print("  step ms:", {k.replace('cng_', ''): round(v, 3) for k, v in r["step_ms_by_entry_point"].items()})
print("  clocks:", d.get("clocks"), " launches:", d.get("gpu_launches"))
if d.get("cpu_baseline"):
    print("  cpu_baseline:", d["cpu_baseline"]["value"], d["cpu_baseline"]["unit"], "on", d["cpu_baseline"]["cores"], "cores")
