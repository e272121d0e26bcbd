#!/usr/bin/env python
"""Hot spots of one ncu --set full capture (source page): kernel duration, instruction totals, the most-sampled SASS
instructions and the sample share per 250-instruction region.  Usage: python tools/ncu_hot.py report.ncu-rep [top]"""
import collections, csv, io, subprocess, sys

def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))

def main():
    rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    raw = page(rep, "raw")
    h, v = raw[0], raw[2]
    for a, b in zip(h, v):
        if a in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
                 "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__cycles_active.avg"):
            print(a, b)
    rows = page(rep, "source")
    h = rows[1]; ix = {n: i for i, n in enumerate(h)}
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    data = []
    for k, r in enumerate(rows[2:]):
        try:
            s = float(r[ix["# Samples"]] or 0)
        except (ValueError, IndexError):
            continue
        data.append((s, k, r))
    tot = sum(d[0] for d in data)
    print("static instructions", len(data), "samples", tot)
    agg = collections.Counter()
    for s, k, r in data:
        for c in stalls:
            try: agg[c] += float(r[ix[c]] or 0)
            except ValueError: pass
    print("stall mix:", [(c, int(n)) for c, n in agg.most_common(7)])
    for s, k, r in sorted(data, key=lambda d: -d[0])[:top]:
        why = max(stalls, key=lambda c: float(r[ix[c]] or 0))
        print(f"{int(s):5d} @{k:5d} exec {r[ix['Instructions Executed']]:>7} thr {r[ix['Avg. Threads Executed']]:>5} {why:18s} {r[ix['Source']][:80]}")
    reg = collections.Counter(); ex = collections.Counter()
    for s, k, r in data:
        reg[k // 250] += s
        try: ex[k // 250] += float(r[ix["Instructions Executed"]] or 0)
        except ValueError: pass
    print("region: samples / executed")
    print("  ".join(f"{k*250}:{int(reg[k])}/{int(ex[k])}" for k in sorted(reg) if ex[k]))

if __name__ == "__main__":
    main()
