"""Markdown summary of bench.py JSON lines (profiles/README.md tables)."""
import json
import sys


def load(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


rows = [load(p) for p in sys.argv[1:]]
print("| N | value (regs/s) | ms/step | e2e (regs/s) | full module (regs/s) | C4 batch (regs/s) | C5 k=1 NCCL / fused (ms) | C5 k=6 NCCL / fused (ms) | C5 map points |")
print("|---|---|---|---|---|---|---|---|---|")
for d in rows:
    b = d.get("batch_lc") or {}
    s = d.get("sharded_knn") or {"cases": [], "map_points": 0}
    c = {x["k"]: x for x in s["cases"]}

    def pair(k):
        if k not in c:
            return "-"
        f = c[k].get("fused_peer_memory", {})
        return "%.3f / %s" % (c[k]["ms"], ("%.3f" % f["ms"]) if "ms" in f else "n/a")
    print("| %d | %.0f | %.3f | %.0f | %.0f | %s | %s | %s | %s |" % (
        d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_full_module"]["registrations_per_s"],
        ("%.0f" % b["registrations_per_s"]) if b else "-", pair(1), pair(6), "{:,}".format(s["map_points"])))
d = rows[0]
print()
print("N=1 details: roofline", json.dumps(d["roofline"]))
print("kernel_ms", json.dumps(d["kernel_ms"]), "gpu_launches", d["gpu_launches"], "clocks", json.dumps(d["clocks"]))
for k in d.get("knn", []):
    print("knn", json.dumps(k))
for key in ("scan_to_map", "accuracy", "e2e_decimated_1m", "cpu_baseline"):
    if d.get(key):
        print(key, json.dumps(d[key]))
