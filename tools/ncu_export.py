"""Development helper: `ncu -i report --page raw --csv` of one captured launch -> the transposed "metric,unit,value" file kept
under profiles/.   usage: ncu_export.py report.ncu-rep out.csv [launch index, default last]"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
names, units, vals = rows[0], rows[1], rows[2:][which]
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for n, u, v in zip(names, units, vals):
        w.writerow([n, u, v])
print(out, len(names), "metrics")
