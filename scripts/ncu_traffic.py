"""Development aid: DRAM bytes per kernel launch from `ncu --set full` reports -> profiles/r2_traffic.json
(what bench.py reports as roofline.traffic).  Usage: ncu_traffic.py key=report.ncu-rep ...   key = "<log2n>/<K>/<block|detect>" """
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "r2_traffic.json")
try:
    out = json.load(open(path))
except Exception:
    out = {}
for arg in sys.argv[1:]:
    key, rep = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    entry = out.setdefault(key, {})
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
        def val(m):
            i = hdr.index(m)
            v = float(r[i].replace(",", ""))
            u = units[i]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
        entry[name] = {"dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
                       "time_ms_under_ncu": val("gpu__time_duration.sum"), "report": os.path.basename(rep)}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
