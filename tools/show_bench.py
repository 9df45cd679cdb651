"""prints the interesting fields of a bench.py JSON line (development aid)"""
import json
import sys

txt = open(sys.argv[1]).read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][0])
print("n_gpus %d  value %.4g  e2e %.4g (packed %.4g)  cpu %s" % (
    d["n_gpus"], d["value"], d["e2e"]["value"], d["e2e"]["packed"]["value"], d.get("cpu_baseline", {}).get("value")))
print("ms_per_step %.4f  frac %.3f  launches %d  ring: %s  clocks %s" % (
    d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["config"]["ring"], d["clocks"]))
print("quality", d.get("quality"))
for k, v in d.get("other_configs", {}).items():
    print(k, "->", {kk: (vv if kk not in ("roofline", "cpu_baseline") else (round(vv.get("frac"), 4) if kk == "roofline" else vv.get("value")))
                    for kk, vv in v.items() if kk not in ("note", "islands_per_gpu", "transport", "unit")})
