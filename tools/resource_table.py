"""ptxas resource table + Blackwell SASS mnemonic census of the shipped kernels (CPU only: nvcc
cross-compiles, cuobjdump disassembles; no GPU needed).

  python tools/resource_table.py            # writes profiles/r02_ptxas_resources.txt

Every source of gae_dgl_b200/build.py::SOURCES is compiled with the build's own flags plus
-Xptxas=-v into a scratch directory (the in-tree library is not touched); the census counts, per
kernel of the in-tree libgae_b200.so, the instructions that prove the sm_100a-specific paths
(/opt/skills/guides/B200_PROFILING.md): UTCMMA / UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), LDGSTS (cp.async), HMMA (legacy mma.sync), MUFU."""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gae_dgl_b200 import build as B  # noqa: E402

CENSUS = ("UTCMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "LDGSTS", "HMMA", "MUFU", "FFMA2")


def demangle(name: str) -> str:
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    return re.sub(r"\(.*$", "", out).replace("void ", "")


def ptxas_rows(scratch: str):
    procs = []
    for s in B.SOURCES:
        cmd = [B._nvcc()] + B.NVCC_FLAGS + ["-Xptxas=-v", "-c", os.path.join(B.CSRC, s), "-o", os.path.join(scratch, s + ".o")]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    rows = []
    for s, p in procs:
        txt, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{txt}")
        parts = re.split(r"ptxas info\s+: Compiling entry function '([^']+)' for 'sm_100a'", txt)
        for i in range(1, len(parts), 2):
            body = parts[i + 1]
            grab = lambda pat: (re.search(pat, body) or [None, "0"])[1]  # noqa: E731
            spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", body)
            rows.append((s, demangle(parts[i]), int(grab(r"Used (\d+) registers")), int(grab(r"used (\d+) barriers")),
                         int(grab(r"(\d+) bytes smem")), "/".join(spill.groups()) if spill else "0/0/0"))
    return rows


def sass_census(lib: str):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = out.setdefault(demangle(m.group(1)), Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for key in CENSUS:
                if op == key or op.startswith(key + "."):
                    cur[key] += 1
                    break
    return out


def main():
    lib = B.build()
    with tempfile.TemporaryDirectory() as scratch:
        rows = ptxas_rows(scratch)
    census = sass_census(lib)
    path = os.path.join(ROOT, "profiles", "r02_ptxas_resources.txt")
    with open(path, "w") as o:
        o.write("# ptxas -v resource usage of every kernel in libgae_b200.so (final round-2 build); tools/resource_table.py\n")
        o.write("# " + " ".join([os.path.basename(B._nvcc())] + B.NVCC_FLAGS + ["-Xptxas=-v"]) + "\n")
        o.write("# smem_B = static shared memory (the tcgen05 / TMA kernels add dynamic shared memory at launch);\n")
        o.write("# spills = stack frame / spill stores / spill loads, bytes\n")
        o.write(f"{'source':16s} {'regs':>4s} {'bar':>3s} {'smem_B':>7s} {'spills':>9s}  kernel\n")
        for r in rows:
            o.write(f"{r[0]:16s} {r[2]:4d} {r[3]:3d} {r[4]:7d} {r[5]:>9s}  {r[1]}\n")
        spilled = [r for r in rows if r[5] != "0/0/0"]
        o.write(f"# {len(rows)} kernels, {len(spilled)} with a stack frame or spills: " + ", ".join(r[1] for r in spilled) + "\n\n")
        o.write("# SASS census of the linked library (cuobjdump -sass), kernels holding at least one of the listed instructions\n")
        o.write("# " + " ".join(CENSUS) + "\n")
        tot = Counter()
        for name in sorted(census):
            c = census[name]
            tot.update(c)
            if any(c[k] for k in CENSUS if k not in ("MUFU", "FFMA2")):
                o.write(f"{name:60s} " + " ".join(f"{k}={c[k]}" for k in CENSUS if c[k]) + "\n")
        o.write("# whole library: " + " ".join(f"{k}={tot[k]}" for k in CENSUS) + "\n")
    print(path)


if __name__ == "__main__":
    main()
