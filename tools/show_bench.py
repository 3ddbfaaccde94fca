#!/usr/bin/env python
"""Compact view of a bench.py JSON line.  Usage: show_bench.py FILE"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g %s  ms/step %.3f  n_gpus %d  e2e %.4g (%.1f ms/step)"
      % (d["value"], d["unit"], d["ms_per_step"], d["n_gpus"],
         d["e2e"]["value"], d["e2e"].get("ms_per_step", 0)))
r = d.get("roofline") or {}
print("roofline: %s frac %.3f" % (r.get("op"), r.get("frac", 0)))
for k, v in (r.get("per_op") or {}).items():
    print("  %-44s %s" % (k, {a: (round(b, 4) if isinstance(b, float) else b)
                             for a, b in v.items()
                             if a in ("ms", "frac", "first_call_ms", "error",
                                      "frac_of_shared_memory_floor")}))
for k in ("no_cache", "onepass_rowVars", "resident"):
    v = d["e2e"].get(k)
    if v:
        print("  e2e.%s: %s" % (k, {a: v[a] for a in ("value", "ms_per_step",
                                                      "error") if a in v}))
for k, v in (d["e2e"].get("per_config") or {}).items():
    print("  e2e %s: %s" % (k, json.dumps(v)[:400]))
cb = d.get("cpu_baseline") or {}
print("cpu_baseline: %s %s cores %s" % (cb.get("value"), cb.get("unit"),
                                        cb.get("cores")))
print("parity:", json.dumps(d["config"].get("parity_checked"))[:900])
print("launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
