#!/usr/bin/env python
"""profiles/rN_*_traffic.json from the ncu summary csv (scripts/ncu_csv.py) of the first flush of `bench.py --frames 4 --steps 1`
(waves: I x64, P x64, B+B x128): DRAM bytes and executed warp instructions per 1080p picture, by kernel and picture type.
bench.py scales them to its workload for `roofline.traffic` and the issue-slot roofline.
Usage: ncu_traffic.py <summary.csv> <source note> > profiles/r2_vNN_traffic.json"""
import csv, json, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
short = {"residual_kernel": "residual", "recon_inter2_kernel": "inter", "recon_intra_kernel": "intra", "recon_intra_sparse_kernel": "intra",
         "deblock_prep_kernel": "deblock_prep", "deblock_kernel": "deblock", "intra_list_kernel": "intra_list"}
# launch order of one flush: per wave the side kernels (list, residual, prep) and the main kernels (inter, intra, deblock)
seen = {}
waves = ["I", "P", "B"]
pics = {"I": 64, "P": 64, "B": 128}
bytes_pp, inst_pp, ms_pp = {}, {}, {}
for r in rows[2:]:
    m = re.search(r"(\w+)\s*(?:<[^(]*>)?\s*\(", r[ix["Kernel Name"]].replace("(bool)", ""))      # template arguments and return type dropped
    name = m.group(1) if m else r[ix["Kernel Name"]].split("(")[0]
    k = short.get(name)
    if not k:
        continue
    # the n-th launch of a kernel belongs to the n-th wave that launches it (the I wave has no inter / sparse / list kernels)
    n = seen.get(name, 0)
    seen[name] = n + 1
    if name in ("recon_inter2_kernel", "recon_intra_sparse_kernel", "intra_list_kernel"):
        wave = waves[1 + n] if 1 + n < 3 else None
    elif name == "recon_intra_kernel":
        wave = "I" if n == 0 else None
    else:
        wave = waves[n] if n < 3 else None
    if wave is None:
        continue
    rd, wr = float(r[ix["dram__bytes_read.sum"]]), float(r[ix["dram__bytes_write.sum"]])
    ru, wu = rows[1][ix["dram__bytes_read.sum"]], rows[1][ix["dram__bytes_write.sum"]]
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = rd * mult[ru] + wr * mult[wu]
    bytes_pp.setdefault(k, {}).setdefault(wave, 0.0)
    bytes_pp[k][wave] += total / pics[wave]
    inst_pp.setdefault(k, {}).setdefault(wave, 0.0)
    inst_pp[k][wave] += float(r[ix["smsp__inst_executed.sum"]]) / pics[wave]
    ms_pp.setdefault(k, {}).setdefault(wave, 0.0)
    ms_pp[k][wave] += float(r[ix["gpu__time_duration.sum"]])
json.dump({"source": sys.argv[2], "unit": "per 1080p picture, by kernel and picture type (I / P / B)",
           "bytes_per_picture": bytes_pp, "warp_inst_per_picture": inst_pp, "ncu_ms_per_wave": ms_pp}, sys.stdout, indent=1)
