"""Static resource table of every kernel in libmvo.so: `nvcc -Xptxas -v` per translation unit (registers, stack, spills,
static shared memory).  Usage: python tools/ptxas_report.py > profiles/ptxas_rN.txt   (no GPU needed)."""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "monocular-visual-odometry_b200" / "csrc"


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", o).split("(")[0].replace("void ", "") for o in out]


def main():
    rows = []
    for cu in sorted(SRC.glob("*.cu")):
        r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-Xptxas", "-v",
                            "-I", str(ROOT / "include"), "-c", str(cu), "-o", "/dev/null"], capture_output=True, text=True)
        text = r.stderr
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?Function properties for \1\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r".*?Used (\d+) registers(?:, used (\d+) barriers)?(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", text, re.S):
            rows.append((cu.name, m.group(1), int(m.group(5)), int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(8) or 0)))
    names = demangle([r[1] for r in rows])
    print("# nvcc 12.9 -O3 -gencode arch=compute_100a,code=sm_100a -Xptxas -v; one row per __global__ entry (template instances included)")
    print(f"{'file':<12} {'kernel':<34} {'regs':>5} {'stack B':>8} {'spill st B':>10} {'spill ld B':>10} {'static smem B':>13}")
    for (f, _, regs, stack, st, ld, smem), n in zip(rows, names):
        print(f"{f:<12} {n:<34} {regs:>5} {stack:>8} {st:>10} {ld:>10} {smem:>13}")


if __name__ == "__main__":
    sys.exit(main())
