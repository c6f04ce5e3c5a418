#!/usr/bin/env python
"""Attribute the SASS of one kernel of the built library to source lines.

  python tools/sass_lines.py 'step_kernelILi0ELi10ELb0ELb1E' [--loops] [--top 40]

Needs the library built with -lineinfo (manytor_b200/build.py does).  Prints, per source line,
how many SASS instructions it owns, split by the innermost loop they sit in (loops = backward
branches), so that "instructions per tile" can be read as: tile-loop body + trip count x inner
loop bodies.  Used to decide where instruction-count work pays.
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "manytor_b200", "lib", "libmanytor_b200.so")


def disassemble(lib: str) -> str:
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, capture_output=True)
        cubins = [f for f in os.listdir(d) if f.endswith(".cubin")]
        assert cubins, "no cubin in " + lib
        return subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubins[0])], check=True,
                              capture_output=True, text=True).stdout


def kernel_text(sass: str, pattern: str):
    lines = sass.splitlines()
    start = None
    for i, ln in enumerate(lines):
        if ln.startswith(".text.") and pattern in ln:
            start = i
            name = ln
            break
    if start is None:
        raise SystemExit(f"no kernel matching {pattern!r}")
    end = len(lines)
    for i in range(start + 1, len(lines)):
        if lines[i].startswith(".text.") or lines[i].startswith(".section"):
            end = i
            break
    return name, lines[start + 1:end]


INSN = re.compile(r"^\s*/\*([0-9a-f]+)\*/\s+(.*?);")
FILE = re.compile(r'//## File "(.*)", line (\d+)')
LABEL = re.compile(r"^(\.L_x_\d+):")


def parse(body):
    """-> list of (addr, text, (file, line)), label -> addr"""
    insns, labels, cur = [], {}, ("?", 0)
    pending = []
    for ln in body:
        m = FILE.search(ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = LABEL.match(ln)
        if m:
            pending.append(m.group(1))
            continue
        m = INSN.match(ln)
        if m:
            addr = int(m.group(1), 16)
            for lb in pending:
                labels[lb] = addr
            pending = []
            insns.append((addr, m.group(2).strip(), cur))
    return insns, labels


def loops_of(insns, labels):
    """backward branches -> list of (head addr, tail addr)"""
    out = []
    for addr, text, _ in insns:
        m = re.search(r"BRA\S*\s+[^`]*`\((\.L_x_\d+)\)", text)
        if m and m.group(1) in labels and labels[m.group(1)] <= addr:
            out.append((labels[m.group(1)], addr))
    return sorted(set(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("pattern")
    ap.add_argument("--lib", default=LIB)
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--ops", action="store_true", help="opcode histogram per loop instead of source lines")
    args = ap.parse_args()
    name, body = kernel_text(disassemble(args.lib), args.pattern)
    insns, labels = parse(body)
    loops = loops_of(insns, labels)
    print(name)
    print(f"{len(insns)} instructions, loops (head..tail, size):")
    for h, t in loops:
        print(f"   {h:#06x}..{t:#06x}  {(t - h) // 16 + 1}")

    def innermost(addr):
        best = None
        for h, t in loops:
            if h <= addr <= t and (best is None or (t - h) < (best[1] - best[0])):
                best = (h, t)
        return best

    per = collections.defaultdict(collections.Counter)
    for addr, text, src in insns:
        lp = innermost(addr)
        key = text.split()[0] if args.ops else f"{src[0]}:{src[1]}"
        if text.startswith("@"):
            key = text.split()[1] if args.ops else key
        per[lp][key] += 1
    for lp in sorted(per, key=lambda k: (-1, -1) if k is None else k):
        tot = sum(per[lp].values())
        print(f"\n== {'outside loops' if lp is None else f'loop {lp[0]:#06x}..{lp[1]:#06x}'}: {tot} instructions")
        for key, c in per[lp].most_common(args.top):
            print(f"   {c:5d}  {key}")


if __name__ == "__main__":
    main()
