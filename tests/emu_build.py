"""TEST INFRASTRUCTURE (CPU tier): host builds of the product's CUDA translation units.

`build(tmp_dir, units)` turns each .cu file into a C++ file the host compiler accepts — kernel launches
`k<<<grid, block, smem, stream>>>(args);` become `emu_launch(grid, block, smem, [&] { k(args); });`, dynamic shared-memory
declarations become pointers into the emulated block's buffer, the one inline-PTX read of %lanemask_lt becomes its
emulation — prepends tests/cpp/cuda_emu.h (one OS thread per CUDA thread) and links the result with
tests/emu/cuda_runtime_emu.cpp (the CUDA runtime calls of the host code over host memory).  Kernel and host code are
otherwise compiled as written.  Used by tests/test_orb_emu.py; never by the product."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "monocular-visual-odometry_b200" / "csrc"

ALL_UNITS = ["ctx.cu", "orb.cu", "orb_host.cpp", "match.cu", "match_host.cpp", "pnp.cu", "ba.cu", "track.cu", "tracker.cpp", "epipolar.cu",
             "two_view.cpp", "motion_host.cpp", "vo_host.cpp", "vo_io.cpp", "vo_pipeline.cpp"]
CLUSTER_UNITS = {"ba.cu"}            # kernels launched as thread-block clusters: their blocks run concurrently in the emulation

LAUNCH = re.compile(r"(\b[A-Za-z_]\w*(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", re.S)
DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];")
DYN_SMEM_MACRO = re.compile(r"(#define\s+\w+\(type, name\))\s+extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?type name\[\]")
LANEMASK = re.compile(r'asm\("mov\.u32 %0, %%lanemask_lt;"\s*:\s*"=r"\((\w+)\)\);')


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


SHARED_DECL = re.compile(r"^(?P<indent>[ \t]*)__shared__\s+(?P<body>[^;=]+);(?P<tail>.*)$", re.M)
DECLARATOR = re.compile(r"^(?P<name>\w+)\s*(?P<dims>(?:\[[^\]]*\])*)$")
_shared_id = [0]


def per_block_shared(text):
    """`__shared__ T a[N], b;` -> references into the emulated block's own storage (clusters run their blocks concurrently)."""
    def repl(m):
        body = m.group("body").strip()
        if body.startswith("extern"):
            return m.group(0)
        body = re.sub(r"__align__\(\d+\)\s*", "", body)
        first = re.search(r"\b\w+\s*(?:\[[^\]]*\])*\s*(?:,|$)", body)
        # the type is everything before the first declarator
        decls = _split_args(body)
        head = decls[0]
        mm = re.match(r"^(?P<type>.+?)\s+(?P<decl>\w+\s*(?:\[[^\]]*\])*)$", head)
        assert mm, body
        ctype = mm.group("type")
        out = []
        for d in [mm.group("decl")] + decls[1:]:
            dm = DECLARATOR.match(d.strip())
            assert dm, (body, d)
            name, dims = dm.group("name"), dm.group("dims")
            i = _shared_id[0]
            _shared_id[0] += 1
            if dims:
                out.append(f"auto &{name} = *reinterpret_cast<{ctype} (*){dims}>(emu_block_static({i}, sizeof({ctype}{dims})));")
            else:
                out.append(f"auto &{name} = *reinterpret_cast<{ctype} *>(emu_block_static({i}, sizeof({ctype})));")
        return m.group("indent") + " ".join(out) + m.group("tail")
    return SHARED_DECL.sub(repl, text)


def transform(text, cluster_kernels=False):
    def launch(m):
        kernel, cfg, args = m.group(1), _split_args(m.group(2)), m.group(3)
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        return f"emu_launch({grid}, {block}, {smem}, [&] {{ {kernel}({args}); }});"
    text, n = LAUNCH.subn(launch, text)
    text = DYN_SMEM.sub(r"\1 *\2 = (\1 *)g_dyn_smem;", text)
    text = DYN_SMEM_MACRO.sub(r"\1 type *name = (type *)g_dyn_smem", text)
    text = LANEMASK.sub(r"\1 = emu_lanemask_lt();", text)
    text = re.sub(r"#include <cooperative_groups.h>\n", "", text)
    text = text.replace("cudaLaunchKernelEx(", "emu_cudaLaunchKernelEx(")
    if cluster_kernels:
        text = per_block_shared(text)                      # `namespace cg = cooperative_groups;` stays: cuda_emu.h provides the namespace
    else:
        text = re.sub(r"namespace cg = cooperative_groups;\n", "", text)
    return '#include "cuda_emu.h"\n' + text, n


def build(tmp_dir, units, name="libmvo_emu.so"):
    """units: file names under csrc/ (.cu are transformed, .cpp compiled as they are).  Returns the path of the shared object."""
    tmp_dir = Path(tmp_dir)
    tmp_dir.mkdir(parents=True, exist_ok=True)
    srcs = []
    for u in units:
        src = CSRC / u
        if u.endswith(".cu"):
            text, _ = transform(src.read_text(), cluster_kernels=(u in CLUSTER_UNITS))
            out = tmp_dir / (u + ".emu.cpp")
            out.write_text(text)
            srcs.append(str(out))
        else:
            srcs.append(str(src))
    so = tmp_dir / name
    # -Bsymbolic: the library's own definitions of the CUDA runtime entry points win over a real libcudart that another test may
    # have brought into the process (torch loads it with RTLD_GLOBAL)
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Wl,-Bsymbolic", "-I", str(ROOT / "include"), "-I", str(CSRC),
                    "-I", str(ROOT / "tests" / "cpp"), "-I", "/usr/local/cuda/include", *srcs, str(ROOT / "tests" / "emu" / "cuda_runtime_emu.cpp"),
                    "-o", str(so)], check=True)
    return so
