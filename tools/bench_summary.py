import json, sys
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    print("ms/frame %.3f  fps %.1f  Mrays/s %.1f  e2e fps %.1f  launches %d" % (d["ms_per_step"], d["fps"], d["value"], d["e2e"]["fps"], d["gpu_launches"]))
    r = d.get("roofline")
    if r:
        print("dominant", r["kernel"], "frac %.3f" % r["frac"], {k: (round(v["ms_per_launch"], 3), v["launches_per_frame"]) for k, v in r["per_kernel"].items()})
        print("traversal", {k: round(v, 2) for k, v in r["traversal"].items()})
    if "cpu_baseline" in d:
        print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
