"""Print the headline fields of a bench.py JSON line (gpurun_out/bench_*.json)."""
import json
import sys

d = json.load(open(sys.argv[1]))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "conv ms", d["roofline"]["kernel_ms_per_step"])
print(d["stage_ms_eager"])
print(" ".join(f"{l['name']}={l['ms']*1e3:.0f}" for l in d["conv_layers"]))
i8 = d.get("int8", {})
for k in ("dynamic_amax", "static_calibration"):
    if k in i8:
        print(k, i8[k].get("frames_per_sec"), i8[k].get("conv_ms_per_step"))
