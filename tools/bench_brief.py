"""Print the headline numbers of a bench.py JSON line (development aid)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
st = d.get("stages", {})
print(json.dumps({"value": round(d["value"]), "ms_per_step": round(d["ms_per_step"], 4), "e2e": round(d["e2e"]["value"]),
                  "stages_ms": {k: round(v["ms_per_step"], 4) for k, v in st.items()}, "roofline_frac": round(d["roofline"]["frac"], 4),
                  "parity_spot": d.get("parity_spot"), "launches": d.get("gpu_launches"), "clocks": d.get("clocks")}))
