"""Summarise an ncu --csv launch list (gpu__time_duration.sum) of bench.py: one decode (or one sub-window of a decode) =
the launches from one FIR / threshold launch (or run of them) up to the next.  Prints the median such group (input
resident in HBM) launch by launch, with `launches.py list.csv K` also one whole step of K sub-windows aggregated by kernel,
and the last host-input step (one screening launch per 64 MiB piece) aggregated by kernel."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hdr]
ki, vi = h.index('Kernel Name'), h.index('Metric Value')
seq = [(r[ki].split('(')[0].replace('void ', ''), float(r[vi].replace(',', ''))) for r in rows[hdr + 2:] if len(r) > vi]
steps, cur = [], []
prev_screen = False
for n, v in seq:
    if 'synth' in n:
        continue
    is_screen = 'fir_screen' in n or 'exact_tiled' in n
    if is_screen and not prev_screen and cur:          # a decode starts with its first FIR / threshold launch
        steps.append(cur)
        cur = []
    prev_screen = is_screen
    cur.append((n, v))
if cur:
    steps.append(cur)
print(f"{len(steps)} decode steps in the list (kernels per step: {[len(x) for x in steps]})")
single = [x for x in steps if sum(1 for n, _ in x if 'fir_screen' in n or 'exact_tiled' in n) == 1 and len(x) > 3]
multi = [x for x in steps if sum(1 for n, _ in x if 'fir_screen' in n or 'exact_tiled' in n) > 1]
if single:
    single.sort(key=lambda x: sum(v for _, v in x))
    dev_step = single[len(single) // 2]                 # median device-resident step
    tot = sum(v for _, v in dev_step)
    print("-- device-resident step (median of %d) --" % len(single))
    for n, v in dev_step:
        print(f"{v / 1000:9.1f} us  {100 * v / tot:5.1f} %  {n}")
    print(f"{tot / 1000:9.1f} us  total (serialised, cold-cache launches under ncu: shares, not absolute times)")
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if K > 1 and len(single) >= K:
        # under ncu the launches are serialised in host order: the K sub-windows of a step follow each other
        order = [x for x in steps if x in single]
        step = order[-K:] if len(order) % K == 0 else order[:K]
        agg = collections.OrderedDict()
        for grp in step:
            for n, v in grp:
                a = agg.setdefault(n, [0, 0.0])
                a[0] += 1
                a[1] += v
        tot = sum(a[1] for a in agg.values())
        print(f"-- one device-resident step = {K} sub-windows, aggregated by kernel --")
        for n, (c, v) in agg.items():
            print(f"{v / 1000:9.1f} us  {100 * v / tot:5.1f} %  x{c:<3d} {n}")
        print(f"{tot / 1000:9.1f} us  total of {sum(a[0] for a in agg.values())} launches (serialised under ncu; live, the tails of sub-windows 1..K-1 overlap the next screens)")
if multi:
    agg = collections.OrderedDict()
    for n, v in multi[0]:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("-- host-input (e2e) step, aggregated (the launch list may be cut by ncu -c) --")
    for n, (c, v) in agg.items():
        print(f"{v / 1000:9.1f} us  {100 * v / tot:5.1f} %  x{c:<3d} {n}")
    print(f"{tot / 1000:9.1f} us  total")
