#!/usr/bin/env python3
"""profiles/traffic.json from an ncu --set full report: dram__bytes_read.sum + dram__bytes_write.sum per launch of the hot-path
kernels (the production instance of each = its longest captured launch), and their sum over the seeding stage, which is what
bench.py reports as roofline.traffic.  usage: traffic_from_ncu.py seed.ncu-rep|seed_raw.csv [more ..] > profiles/traffic.json"""
import csv, json, subprocess, sys

SEEDING = ("pack_reads_kernel", "smem_sweep_kernel", "smem_fast_p3_kernel")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {"source": "ncu --set full --clock-control none (tools/ncu_round.sh), reports: " + ", ".join(sys.argv[1:]), "kernels": {}}
for rep in sys.argv[1:]:
    txt = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("bwaflow::", "")
        t = float(r[col["gpu__time_duration.sum"]])
        rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[units[col["dram__bytes_write.sum"]]]
        k = out["kernels"].get(name)
        if k is None or t > k["time"]:
            out["kernels"][name] = {"time": t, "time_unit": units[col["gpu__time_duration.sum"]], "dram_read_bytes": rd, "dram_write_bytes": wr,
                                    "l2_hit_pct": float(r[col["lts__t_sector_hit_rate.pct"]])}
out["seeding_stage"] = sum(v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in out["kernels"].items() if k.startswith(SEEDING))
json.dump(out, sys.stdout, indent=1)
