"""Reads `ncu -i <rep> --page raw --csv` output on stdin and writes the DRAM traffic of the captured kernel into
profiles/score_traffic.json (the `roofline.traffic` figure of bench.py): python tools/ncu_traffic.py <ctx> <rep-name>"""
import csv
import json
import os
import sys

ctx, rep = sys.argv[1], sys.argv[2]
rows = list(csv.reader(sys.stdin))
h = rows[0]
units = rows[1]
v = rows[2]
get = lambda name: (float(v[h.index(name)].replace(",", "")), units[h.index(name)])
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd, ru = get("dram__bytes_read.sum")
wr, wu = get("dram__bytes_write.sum")
dur, du = get("gpu__time_duration.sum")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "score_traffic.json")
data = json.load(open(path)) if os.path.exists(path) else {}
data[str(ctx)] = int(rd * scale[ru] + wr * scale[wu])
data.setdefault("_provenance", {})[str(ctx)] = {"ncu_report": rep, "dram_read_bytes": int(rd * scale[ru]), "dram_write_bytes": int(wr * scale[wu]),
                                                "gpu_time_duration": f"{dur} {du}", "written_by": "tools/ncu_traffic.py"}
json.dump(data, open(path, "w"), indent=1)
print(json.dumps(data[str(ctx)]), data["_provenance"][str(ctx)])
