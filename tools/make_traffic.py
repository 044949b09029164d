"""profiles/traffic.json (DRAM bytes per launch of each fused kernel, read by bench.py's roofline leg)
from the committed ncu summary CSV of the CURRENT code:

  gpurun: ncu --set full --clock-control none --import-source on -k regex:xview -o gpurun_out/prof_r2_xview \
              python tools/profile_kernels.py
  here  : python tools/summarize_profiles.py full gpurun_out/prof_r2_xview.ncu-rep profiles/r2_xview_kernels_ncu_full.csv \
              Cw_f32_N6_fwd Cw_f32_N6_bwd ... (the launch order of tools/profile_kernels.py, REPS times each)
          python tools/make_traffic.py profiles/r2_xview_kernels_ncu_full.csv
"""
import collections
import csv
import json
import os
import sys

KEYS = {"Cw_f32_N6": "N6_f32", "Cw_f32_N12": "N12_f32", "Cw_bf16_N12": "N12_bf16", "Cn_f32_N12": "N12_f32_Cn",
        "A_f32_N6": "N6_f32_A"}


def main(path):
    rows = list(csv.DictReader(open(path)))
    rd = next(k for k in rows[0] if k.startswith("dram__bytes_read.sum"))
    wr = next(k for k in rows[0] if k.startswith("dram__bytes_write.sum"))
    scale = {"[byte]": 1, "[Kbyte]": 1e3, "[Mbyte]": 1e6, "[Gbyte]": 1e9}
    slots = collections.OrderedDict()                       # a sorted backward call is five kernels: sum its slot
    for r in rows:
        b = float(r[rd]) * scale[rd.split()[-1]] + float(r[wr]) * scale[wr.split()[-1]]
        key = (r["case"], r.get("slot", len(slots)))
        slots[key] = slots.get(key, 0.0) + b
    acc = collections.defaultdict(list)
    for (case, _), b in slots.items():
        if case:
            acc[case].append(b)
    for r in rows:                                           # the owner kernel of a sorted backward, on its own
        if r["case"].endswith("_bwdS") and "xview_bwd_owner_kernel" in r[next(k for k in r if k.startswith("Kernel Name"))]:
            acc[r["case"] + "owner"].append(float(r[rd]) * scale[rd.split()[-1]] + float(r[wr]) * scale[wr.split()[-1]])
    out = collections.defaultdict(dict)
    for case, vals in acc.items():
        base, direction = case.rsplit("_", 1)
        out[KEYS[base]][direction] = sum(vals) / len(vals)
    out["_source"] = os.path.basename(path)
    dst = os.path.join(os.path.dirname(os.path.abspath(path)), "traffic.json")
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
