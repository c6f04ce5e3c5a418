#!/usr/bin/env python
"""Static check for the hazard that cost 21 % of the step kernel's stall samples in round 1: a prefetch
load (LDG / LDG.128 issued one tile ahead) whose destination register is WRITTEN by a later
instruction before anything has READ it.  The overwrite has to wait for the load to land
(write-after-write on the scoreboard), so the prefetch turns into a blocking load.
  cuobjdump -sass lib.so | python tools/sass_waw.py [kernel-name-substring]
Linear scan in address order from each LDG up to the next unconditional branch; predicated code is
treated as straight-line, so a report is a pointer to look at, not a proof."""
import re, sys

INS = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);")
REG = re.compile(r"\bR(\d+)\b")
WIDE = {"128": 4, "64": 2}


def dest_regs(text):
    body = text.split(None, 1)
    if text.startswith("@"):
        body = body[1].split(None, 1)
    if len(body) < 2:
        return set(), set(), body[0]
    op, args = body
    parts = [a.strip() for a in args.split(",")]
    if op.split(".")[0] in ("STG", "STS", "ST", "BRA", "EXIT", "BSSY", "BSYNC", "SYNCS", "UBLKCP", "ATOMG", "RED", "REDG",
                            "ISETP", "FSETP", "PLOP3", "UISETP", "NOP", "YIELD", "WARPSYNC", "BAR", "R2UR", "VOTEU"):
        return set(), {int(m) for m in REG.findall(args)}, op
    d = REG.match(parts[0])
    dst = set()
    if d:
        base = int(d.group(1))
        n = 1
        for k, v in WIDE.items():
            if "." + k in op and op.startswith(("LD", "LDG", "LDS")):
                n = v
        if op.split(".")[0] in ("FFMA2", "FADD2", "FMUL2", "IMAD.WIDE", "CS2R", "DADD", "DFMA") or "WIDE" in op or op.startswith("CS2R") or ".64" in op and op.startswith(("LD", "MOV")):
            n = max(n, 2)
        dst = {base + i for i in range(n)}
    src = set()
    for a in parts[1:] if d else parts:
        for m in REG.finditer(a):
            r = int(m.group(1))
            src.add(r)
            if ".64" in a or "F32x2" in a:
                src.add(r + 1)
    return dst, src, op


def complementary(a, b):
    pa, pb = re.match(r"@(!?)(U?P\d+)\s", a), re.match(r"@(!?)(U?P\d+)\s", b)
    return bool(pa and pb and pa.group(2) == pb.group(2) and pa.group(1) != pb.group(1))


def main():
    want = sys.argv[1] if len(sys.argv) > 1 else ""
    kern, ins, bad = None, [], 0
    def flush():
        nonlocal bad
        if kern is None or want not in kern:
            return
        for i, (addr, text) in enumerate(ins):
            dst, _, op = dest_regs(text)
            if not op.startswith("LDG"):
                continue
            pending = set(dst)
            for addr2, text2 in ins[i + 1:]:
                d2, s2, op2 = dest_regs(text2)
                if op2.split(".")[0] in ("BRA", "EXIT", "RET") and not text2.startswith("@"):
                    break                          # end of the straight-line region this load belongs to
                pending -= s2                      # read: the scoreboard wait is a true dependency, fine
                if complementary(text, text2):     # "@!P2 LDG" / "@P2 MOV": the two never both execute
                    continue
                hit = pending & d2
                if hit:
                    print(f"{kern[:70]}: {text[:50]} @{addr}: R{sorted(hit)} overwritten unread by '{text2[:60]}' @{addr2}")
                    bad += 1
                    break
                if not pending:
                    break
    for line in sys.stdin:
        if "Function :" in line:
            flush()
            kern, ins = line.split("Function :")[1].strip(), []
            continue
        m = INS.search(line)
        if m and not m.group(2).startswith("/*"):
            ins.append((m.group(1), m.group(2).strip()))
    flush()
    print(f"{bad} suspicious prefetch overwrite(s)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
