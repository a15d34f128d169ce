"""tools/sass_ctrl.py <cuobjdump -sass output> [pattern] [before] [after]: SASS with the decoded scheduling control fields
(stall count, write / read barrier index, wait mask over the six scoreboards) — what a 'long scoreboard' stall sample of
ncu's source page is actually waiting for."""
import re
import sys


def decode(path):
    lines = open(path).read().split("\n")
    pat = re.compile(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
    hexp = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
    out, i = [], 0
    while i < len(lines):
        m = pat.search(lines[i])
        if m and i + 1 < len(lines):
            m2 = hexp.match(lines[i + 1])
            if m2:
                ctrl = (int(m2.group(1), 16) >> 41) & 0x7FFFFF
                out.append((m.group(1), m.group(2).strip(), ctrl & 0xF, (ctrl >> 4) & 1, (ctrl >> 5) & 7, (ctrl >> 8) & 7, (ctrl >> 11) & 0x3F))
                i += 2
                continue
        i += 1
    return out


def fmt(o):
    w = "".join(str(b) if (o[6] >> b) & 1 else "-" for b in range(6))
    return f"{o[0]} st{o[2]:2d} w{o[4] if o[4] != 7 else '-'} r{o[5] if o[5] != 7 else '-'} wait[{w}] {o[1][:100]}"


if __name__ == "__main__":
    out = decode(sys.argv[1])
    if len(sys.argv) > 2:
        before = int(sys.argv[3]) if len(sys.argv) > 3 else 40
        after = int(sys.argv[4]) if len(sys.argv) > 4 else 20
        for k, o in enumerate(out):
            if sys.argv[2] in o[1]:
                print(f"---- match at {k}")
                for p in out[max(0, k - before):k + after]:
                    print(fmt(p))
    else:
        for o in out:
            print(fmt(o))
