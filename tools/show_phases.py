"""Print the per-phase k-cycles of bench lines: python tools/show_phases.py gpurun_out/var_a.json [gpurun_out/var_b.json ...]"""
import json
import sys

rows = []
for f in sys.argv[1:]:
    try:
        d = json.load(open(f))
    except Exception as e:  # noqa: BLE001
        print(f, "unreadable:", e)
        continue
    rows.append((f, d))
keys = []
for _, d in rows:
    for k in d["roofline"].get("phase_kcycles_per_particle", {}):
        if k not in keys:
            keys.append(k)
print("%-28s" % "phase", *["%12s" % f.split("/")[-1].replace("var_", "").replace(".json", "")[:12] for f, _ in rows])
print("%-28s" % "ms/step", *["%12.3f" % d["ms_per_step"] for _, d in rows])
for k in keys:
    print("%-28s" % k, *["%12.1f" % d["roofline"]["phase_kcycles_per_particle"].get(k, 0.0) for _, d in rows])
