"""Print the kernels of the last bench step from an ncu --csv launch list (gpu__time_duration.sum)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hdr]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
seq = [(r[ki][:60], float(r[vi].replace(',', ''))) for r in rows[hdr + 2:]]
idx = [i for i, (n, v) in enumerate(seq) if 'screen' in n]
tot = 0
for n, v in seq[idx[-1]:]:
    print(f"{v / 1000:9.1f} us  {n}"); tot += v
print(f"{tot / 1000:9.1f} us  total")
