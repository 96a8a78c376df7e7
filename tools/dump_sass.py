"""Regenerate profiles/sass/: one cuobjdump listing per kernel + an INDEX of the
mnemonics that prove which hardware path each kernel uses (tcgen05 = UTCHMMA /
LDTM, TMA = UTMALDG / UTMASTG / UBLKCP, NVLS = LDGMC / ...MULTIMEM..., sys-scope
release/acquire = *.STRONG.SYS).  Runs on the CPU box (no GPU needed)."""
from __future__ import annotations

import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "sass")
INTEREST = re.compile(
    r"^(LDG|STG|LDGMC|RED|ATOMG|UTC|UTMA|UBLKCP|LDTM|STTM|SYNCS|MEMBAR|ERRBAR|CCTL|REDUX|MULTIMEM|UTMAPF|FENCE)"
)


def main() -> int:
    sos = glob.glob(os.path.join(ROOT, "rayfed_b200", "_fedcuda*.so"))
    if not sos:
        print("build first: python __graft_entry__.py", file=sys.stderr)
        return 1
    txt = subprocess.run(
        ["cuobjdump", "-sass", sos[0]], check=True, capture_output=True, text=True
    ).stdout
    os.makedirs(OUT, exist_ok=True)
    for old in glob.glob(os.path.join(OUT, "*.sass")):
        os.remove(old)
    # split per function
    chunks = re.split(r"^\s*Function : ", txt, flags=re.M)[1:]
    index = [
        "# SASS listings (cuobjdump -sass of rayfed_b200/_fedcuda*.so, sm_100a)",
        "",
        "Regenerate with `python tools/dump_sass.py`.  Encoding lines stripped.",
        "Mnemonic counts of interest per kernel:",
        "",
    ]
    for ch in chunks:
        mangled, _, body = ch.partition("\n")
        mangled = mangled.strip()
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip() or mangled
        short = re.sub(r"[^A-Za-z0-9_]+", "_", dem.split("(")[0].replace("void ", "").replace("fedk::", "")).strip("_")
        lines = []
        counts = collections.Counter()
        for ln in body.splitlines():
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", ln)
            if not m:
                continue
            ins = m.group(1)
            lines.append(ins)
            op = re.sub(r"^@!?U?P\d+\s+", "", ins).split(" ")[0].rstrip(";")
            if INTEREST.match(op):
                counts[op] += 1
        fn = f"{short}.sass"
        with open(os.path.join(OUT, fn), "w") as f:
            f.write(f"// {dem}\n" + "\n".join(lines) + "\n")
        top = dict(counts.most_common(16))
        index.append(f"* `{dem}` → `{fn}`  \n  {top}")
    # static resource usage (cuobjdump -res-usage): registers / static smem / stack / local per kernel
    res = subprocess.run(["cuobjdump", "-res-usage", sos[0]], check=True, capture_output=True, text=True).stdout
    rows = []
    cur = None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            continue
        if cur and "REG:" in ln:
            kv = dict(re.findall(r"(REG|STACK|SHARED|LOCAL|CONSTANT\[0\]):(\d+)", ln))
            rows.append((cur.replace("void ", "").replace("fedk::", ""), kv))
            cur = None
    index += ["", "## Static resources per kernel (`cuobjdump -res-usage`; dynamic shared memory is set at launch)", "",
              "| kernel | registers/thread | static smem (B) | stack (B) | local (B) |", "|---|---:|---:|---:|---:|"]
    for name, kv in rows:
        index.append(f"| `{name}` | {kv.get('REG', '?')} | {kv.get('SHARED', '?')} | {kv.get('STACK', '?')} | "
                     f"{kv.get('LOCAL', '?')} |")
    with open(os.path.join(OUT, "INDEX.md"), "w") as f:
        f.write("\n".join(index) + "\n")
    print(f"{len(chunks)} kernels → {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
