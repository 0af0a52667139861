"""Instruction mix of the innermost loops of a kernel in libghostcwt.so (cuobjdump -sass)."""
import collections
import re
import subprocess
import sys


def main():
    pat = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else "ghost_b200/libghostcwt.so"
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", txt)
    for b in blocks:
        name = b.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for l in b.split("\n"):
            m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        loops = []
        for addr, t in ins:
            m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < addr:
                loops.append((int(m.group(1), 16), addr))
        print(name[:100], "total", len(ins))
        for lo, hi in loops:
            body = [t for a, t in ins if lo <= a <= hi]
            if len(body) < 200:
                continue
            c = collections.Counter()
            for t in body:
                t = re.sub(r"^@!?U?P\d+\s+", "", t)
                c[t.split()[0].split(".")[0]] += 1
            fp = sum(c[k] for k in ("FADD", "FFMA", "FMUL"))
            print("  loop %#x-%#x: %d instrs (%.1f per output), fp %d" % (lo, hi, len(body), len(body) / 16, fp))
            print("   ", ", ".join("%s %d" % kv for kv in c.most_common(14)))


if __name__ == "__main__":
    main()
