"""Condense a bench.py JSON line (stdin) to the few numbers the kernel sweeps compare."""
import json, sys
for l in sys.stdin:
    l = l.strip()
    if not l.startswith("{"):
        continue
    d = json.loads(l)
    st = d.get("stages", {})
    print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]),
          " ".join("%s %.3f" % (k, v["ms"]) for k, v in st.items()),
          "tensorTF", st.get("coarse", {}).get("tensor_TFLOPs_executed"),
          "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
