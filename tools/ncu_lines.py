#!/usr/bin/env python
"""Warp-stall samples of an ncu source-page CSV (`ncu -i rep --page source --csv`, SASS view) aggregated per CUDA SOURCE
LINE: the k-th SASS instruction of the kernel is looked up in `nvdisasm -g` of the same build's cubin (-lineinfo).
Usage: python tools/ncu_lines.py <src.csv> <kernel substring> [cubin] [top]
       cubin defaults to the one extracted from libmvo.so whose text holds the kernel."""
import collections, csv, re, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent

def cubins():
    d = Path(tempfile.mkdtemp(prefix="mvo_cub_"))
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "monocular-visual-odometry_b200" / "libmvo.so")], cwd=d, capture_output=True)
    return sorted(d.glob("*.cubin"))

def line_table(cubin, kernel):
    """[(file, line)] for every instruction of the first entry function whose mangled name contains `kernel`."""
    out = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
    tab, inside, cur = [], False, ("?", 0)
    for l in out:
        if l.startswith(".text.") and l.rstrip().endswith(":"):
            if inside and tab:
                break
            inside = kernel in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            tab.append(cur)
    return tab

def main():
    src, kernel = sys.argv[1], sys.argv[2]
    cub = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3].endswith(".cubin") else None
    top = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 25
    tab = None
    for c in ([Path(cub)] if cub else cubins()):
        tab = line_table(c, kernel)
        if tab:
            break
    rows = list(csv.reader(open(src)))
    h = rows[1]; ix = {n: i for i, n in enumerate(h)}
    stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    per = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
    n = 0
    for k, r in enumerate(rows[2:]):
        try:
            s = float(r[ix["# Samples"]] or 0); ex = float(r[ix["Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        n += 1
        key = tab[k] if tab and k < len(tab) else ("?", 0)
        per[key][0] += s; per[key][1] += ex
        for c in stalls:
            try: per[key][2][c] += float(r[ix[c]] or 0)
            except ValueError: pass
    tot = sum(v[0] for v in per.values()) or 1
    print(f"{n} SASS instructions ({len(tab or [])} in the cubin), {int(tot)} samples")
    for key, (s, ex, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / tot:5.1f}%  {key[0]}:{key[1]:<5d} exec {int(ex):>8}  " + " ".join(f"{c[6:]}={int(v)}" for c, v in st.most_common(3) if v))

if __name__ == "__main__":
    main()
