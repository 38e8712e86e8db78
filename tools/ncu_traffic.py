"""DRAM bytes per launch of the dominant kernels from this round's `ncu --set full` captures -> profiles/r2_traffic.json
(read by bench.py for `roofline*.traffic`).  Usage: python tools/ncu_traffic.py key=report.ncu-rep[:kernel-substring] ...
dram_bytes = dram__bytes_read.sum + dram__bytes_write.sum of the selected launch (the last one matching the filter)."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out_path = os.path.join(ROOT, "profiles", "r2_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=", 1)
    rep, _, kfilter = rest.partition(":")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kcol = hdr.index("Kernel Name")
    sel = [r for r in rows[2:] if not kfilter or kfilter in r[kcol]]
    r = sel[-1]
    def val(name):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * UNIT[units[i]]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    out[key] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "kernel": r[kcol][:100],
                "duration_us_under_ncu": val("gpu__time_duration.sum") / 1e3 if units[hdr.index("gpu__time_duration.sum")] == "nsecond" else None,
                "source": "profiles/r2_traffic.json <- " + os.path.basename(rep) + " (ncu --set full --clock-control none, one launch)"}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
