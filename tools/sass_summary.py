#!/usr/bin/env python
"""SASS evidence for the hot kernels (runs anywhere cuobjdump is -- no GPU needed): opcode histogram of the whole
kernel and the instructions of its hottest loop (the longest backward-branch span), written under profiles/.

    python tools/sass_summary.py allpairs_fast_kernelILi4ELi2ELi0E profiles/r02_sass_allpairs_fast.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rust_exp_b200", "libnbody_b200.so")


def main():
    pat, out = sys.argv[1], sys.argv[2]
    txt = subprocess.check_output(["cuobjdump", "-sass", SO], text=True)
    name, ins, arch = None, [], None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name and pat in name and ins:
                break
            name, ins = m.group(1), []
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and name and pat in name:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    if not (name and pat in name and ins):
        raise SystemExit(f"no kernel matching {pat}")
    ops = collections.Counter()
    for _, t in ins:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        ops[t.split()[0].split(".")[0] if not t.startswith("{") else "?"] += 1
    # hottest loop: the backward branch with the longest span
    best = (0, 0, 0)
    for addr, t in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr and addr - tgt > best[0]:
                best = (addr - tgt, tgt, addr)
    lines = [f"# {name}", f"# arch: {arch}; {len(ins)} instructions; library: rust_exp_b200/libnbody_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a)",
             "# opcode histogram (whole kernel):"]
    lines += [f"#   {n:6d}  {op}" for op, n in ops.most_common()]
    loop = [(a, t) for a, t in ins if best[1] <= a <= best[2]]
    lops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in loop)
    lines += [f"# hottest loop: 0x{best[1]:04x}..0x{best[2]:04x}, {len(loop)} instructions: " + ", ".join(f"{op} x{n}" for op, n in lops.most_common(14))]
    lines += [f"/*{a:04x}*/  {t} ;" for a, t in loop]
    open(os.path.join(ROOT, out), "w").write("\n".join(lines) + "\n")
    print(f"wrote {out}: {len(ins)} instructions, loop of {len(loop)}")


if __name__ == "__main__":
    main()
