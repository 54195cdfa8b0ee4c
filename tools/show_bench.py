import json, sys
for ln in sys.stdin.read().strip().splitlines():
    try: d = json.loads(ln)
    except Exception: continue
    r = d.get("roofline") or {}
    print(d.get("metric"), "value %.1f M" % (d["value"] / 1e6), "ms/step %.1f" % d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "launches", d.get("gpu_launches"),
          "| roof", r.get("kernel"), "frac %.3f" % (r.get("frac") or 0), "ach %.0f" % (r.get("achieved") or 0))
    for k, v in sorted((r.get("kernel_share_of_step") or {}).items(), key=lambda x: -x[1]): print("   %-28s %.3f" % (k, v))
