#!/usr/bin/env python
"""Static look at a kernel's SASS: decode the scheduling control bits of every instruction (stall count,
yield, write/read barrier, wait mask -- bits 105..125 of the 128-bit encoding, Volta and later) and print
the hottest loop (the innermost backward branch whose body holds the most DMMA/DFMA) with per-instruction
stall counts and their sum = the minimum issue time of one warp per iteration, scoreboard waits excluded.

    python profiles/sass_loop.py <substring of the mangled kernel name> [--so bodge_b200/libbdg.so] [--all]

Runs in the build container (cuobjdump only, no GPU).
"""
import re
import subprocess
import sys


def kernels(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    cur, table = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            table[cur] = []
        elif cur is not None:
            table[cur].append(line)
    return table


def parse(lines):
    """-> list of (addr, text, enc_lo, enc_hi)"""
    ins = []
    i = 0
    while i < len(lines):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*/\* (0x[0-9a-f]{16}) \*/", lines[i])
        if m and i + 1 < len(lines):
            m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                ins.append((int(m.group(1), 16), m.group(2).rstrip(" ;"), int(m.group(3), 16), int(m2.group(1), 16)))
                i += 2
                continue
        i += 1
    return ins


def control(hi):
    c = hi >> 41  # bits 105.. of the 128-bit word = bits 41.. of the high half
    return dict(stall=c & 0xF, yld=(c >> 4) & 1, wbar=(c >> 5) & 7, rbar=(c >> 8) & 7, wait=(c >> 11) & 0x3F)


def loops(ins):
    by_addr = {a: k for k, (a, *_rest) in enumerate(ins)}
    found = []
    for k, (a, text, lo, hi) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in by_addr:
                found.append((by_addr[tgt], k))
    return found


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    so = "bodge_b200/libbdg.so"
    if "--so" in sys.argv:
        so = sys.argv[sys.argv.index("--so") + 1]
        args = [a for a in args if a != so]
    want = args[0]
    table = kernels(so)
    names = [n for n in table if want in n]
    if not names:
        sys.exit(f"no kernel matching {want!r}; have e.g. {list(table)[:5]}")
    for name in names[: (None if "--every" in sys.argv else 1)]:
        ins = parse(table[name])
        print(f"== {name}: {len(ins)} instructions")
        ls = loops(ins)
        score = lambda be: sum(1 for k in range(be[0], be[1] + 1) if re.search(r"DMMA|DFMA|DADD|DMUL", ins[k][1]))
        # innermost = no other loop strictly inside
        inner = [l for l in ls if not any(o != l and o[0] >= l[0] and o[1] <= l[1] for o in ls)]
        cand = sorted(inner if "--outer" not in sys.argv else ls, key=score, reverse=True)
        if not cand:
            print("no loop found")
            continue
        show = cand if "--all" in sys.argv else cand[:1]
        for b, e in show:
            body = ins[b : e + 1]
            total = 0
            hist = {}
            print(f"-- loop {ins[b][0]:#x}..{ins[e][0]:#x}: {len(body)} instructions")
            for a, text, lo, hi in body:
                c = control(hi)
                total += max(c["stall"], 1)
                op = text.split()[1] if text.startswith("@") else text.split()[0]
                op = op.split(".")[0]
                hist[op] = hist.get(op, 0) + 1
                if "--quiet" not in sys.argv:
                    print(f"{a:06x} st={c['stall']:2d} y={c['yld']} w={c['wbar']} r={c['rbar']} wait={c['wait']:06b}  {text}")
            print(f"-- sum of stall counts: {total} cycles / iteration / warp;  opcode histogram: "
                  + ", ".join(f"{k}:{v}" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
