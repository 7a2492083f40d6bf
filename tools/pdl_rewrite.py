#!/usr/bin/env python
"""One-off source rewrite (kept for the record): turn every `kern<<<grid, block, smem, stream>>>(args)` launch into
`ts::launch_k(kern, grid, block, smem, stream, args)` (programmatic dependent launch attribute, see common.cuh) and make every
__global__ function start with `ts::pdl_enter();` (griddepcontrol.wait + launch_dependents). Idempotent."""
import re
import sys


def match_back_angle(s, i):
    """s[i] == '>' : index of the matching '<' going backwards."""
    depth = 0
    while i >= 0:
        if s[i] == '>':
            depth += 1
        elif s[i] == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced <>")


def match_fwd(s, i, o, c):
    depth = 0
    while i < len(s):
        if s[i] == o:
            depth += 1
        elif s[i] == c:
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced " + o + c)


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def rewrite_launches(s):
    n = 0
    while True:
        i = s.find("<<<")
        if i < 0:
            break
        # kernel expression, backwards
        j = i - 1
        while s[j].isspace():
            j -= 1
        if s[j] == '>':
            j = match_back_angle(s, j) - 1
        while j >= 0 and (s[j].isalnum() or s[j] in "_:"):
            j -= 1
        k0 = j + 1
        kern = s[k0:i].strip()
        e = s.find(">>>", i)
        cfg = [c.strip() for c in split_top(s[i + 3:e])]
        while len(cfg) < 4:
            cfg.append("0")
        p0 = e + 3
        while s[p0].isspace():
            p0 += 1
        assert s[p0] == "(", (kern, s[p0:p0 + 20])
        p1 = match_fwd(s, p0, "(", ")")
        args = s[p0 + 1:p1]
        new = f"ts::launch_k({kern}, {cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}" + (", " + args.lstrip() if args.strip() else "") + ")"
        s = s[:k0] + new + s[p1 + 1:]
        n += 1
    return s, n


def add_entry(s):
    n = 0
    pos = 0
    while True:
        m = re.compile(r"__global__").search(s, pos)
        if not m:
            break
        p0 = s.find("(", m.end())
        # skip __launch_bounds__(...) groups
        while True:
            head = s[m.end():p0]
            if "__launch_bounds__" in head and head.rstrip().endswith("__launch_bounds__"):
                p0 = s.find("(", match_fwd(s, p0, "(", ")") + 1)
            else:
                break
        p1 = match_fwd(s, p0, "(", ")")
        b = p1 + 1
        while s[b].isspace():
            b += 1
        if s[b] != "{":      # a declaration
            pos = b
            continue
        after = s[b + 1:b + 60]
        if "pdl_enter" not in after:
            s = s[:b + 1] + "\n  ts::pdl_enter();" + s[b + 1:]
            n += 1
        pos = b + 1
    return s, n


for path in sys.argv[1:]:
    src = open(path).read()
    out, a = rewrite_launches(src)
    out, b = add_entry(out)
    if out != src:
        open(path, "w").write(out)
    print(f"{path}: {a} launches, {b} kernels")
