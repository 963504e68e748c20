"""DRAM traffic per launch of the captured kernels -> profiles/r02_traffic.json (read by bench.py: roofline.traffic).
usage: python tools/ncu_traffic.py c2=rep1.ncu-rep c4=rep2.ncu-rep ... > profiles/r02_traffic.json"""
import csv, io, json, subprocess, sys

out = {"note": "dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the configuration's dominant kernel, from ncu --set full --clock-control none "
               "captures of `python bench.py --config <c> --steps 3 --warmup 3 --no-cpu` (summaries: profiles/r02_*_ncu_full.txt)"}
for arg in sys.argv[1:]:
    cfg, rep = arg.split("=")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    m, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))

    def val(key):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]
        return float(m[key]) * scale

    out[cfg] = {"kernel": m.get("Kernel Name", "?"), "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                "gpu_time_us": float(m["gpu__time_duration.sum"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[u["gpu__time_duration.sum"]]}
print(json.dumps(out, indent=1))
