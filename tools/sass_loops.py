"""List the loops (backward branches) of one kernel in an object file with their SASS instruction counts.

    python tools/sass_loops.py /tmp/xp_fast.o suite_fast_kernelILj7ELi1ELi512ELi0 [--min 100] [--dump-largest FILE]
"""
import re
import subprocess
import sys
from collections import Counter


def main():
    obj, key = sys.argv[1], sys.argv[2]
    mn = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 100
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0]
        if key not in name:
            continue
        ins = []
        for l in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        idx = {a: i for i, (a, _) in enumerate(ins)}
        print(name[:90], len(ins), "SASS instructions")
        for i, (a, s) in enumerate(ins):
            if "BRA" in s:
                t = re.search(r"0x([0-9a-f]+)", s)
                if t and int(t.group(1), 16) < a and int(t.group(1), 16) in idx:
                    j = idx[int(t.group(1), 16)]
                    n = i - j + 1
                    if n >= mn:
                        body = [x[1] for x in ins[j:i + 1]]
                        c = Counter((b.split()[1] if b.startswith("@") else b.split()[0]).split(".")[0] for b in body)
                        print(f"  loop {hex(ins[j][0])}..{hex(a)}: {n} instr", dict(c.most_common(14)))


if __name__ == "__main__":
    main()
