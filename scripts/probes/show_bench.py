import json, sys
d = json.loads(open(sys.argv[1]).read())
print("ms/step %.2f value %.1f M e2e %.1f M" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6), {k: round(v, 2) for k, v in d["roofline"]["kernel_ms_per_step"].items()}, "parity", d["parity_checked"])
