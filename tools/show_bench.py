"""Print the headline fields of bench.py JSON lines (python tools/show_bench.py file...)."""
import json
import sys
for path in sys.argv[1:]:
    lines = [l for l in open(path) if l.startswith("{")]
    if not lines:
        print(path, "no JSON line"); continue
    d = json.loads(lines[-1])
    print("== %s: N=%s  %.1f %s  %.4f ms/step  launches %s  e2e %.2f" % (
        path, d.get("n_gpus"), d["value"], d["unit"], d["ms_per_step"], d.get("gpu_launches"), d["e2e"]["value"]))
    r = d.get("roofline", {})
    if r:
        print("   roofline: bound %s  %.0f GB/s = %.3f of %.0f; fp32 frac %.3f; kernel_ms %.4f sustained %s" % (
            r.get("bound"), r["achieved"], r["frac"], r["peak"], r.get("fp32", {}).get("frac", float("nan")), r["kernel_ms"],
            (r.get("kernel_ms_sustained_blocks") or {}).get("median")))
    print("   parity:", d.get("parity"))
    print("   clocks:", d.get("clocks"))
    if "e2e" in d and "copy_only" in d["e2e"]:
        print("   e2e copy-only bound:", d["e2e"]["copy_only"])
    for k, v in (d.get("legs") or {}).items():
        if "skipped" in v:
            print("   %s skipped: %s" % (k, v["skipped"])); continue
        if "axis1" in v:
            print("   %s: axis1 %.4f ms (%.0f Gpx/s)  axis2 %.4f ms (%.0f Gpx/s)  %s" % (
                k, v["axis1"]["ms_per_step"], v["axis1"]["value"], v["axis2"]["ms_per_step"], v["axis2"]["value"], v["parity"]))
        else:
            print("   %s: %.4f ms  %.0f Gvoxel/s  launches/gpu %s  %s" % (k, v["ms_per_step"], v["value"], v["launches_per_step_per_gpu"], v["halo_backend"][:60]))
            print("        ", {a: b for a, b in v["parity"].items() if a not in ("tol", "brick")})
