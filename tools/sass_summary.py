"""Static evidence of the built library (no GPU needed): per kernel registers / shared memory / spills from the
ptxas logs of `make lib` (build/obj/*.ptxas.log), and the SASS mnemonics that show what the kernels are made of
(UTMALDG / UTMASTG = TMA tile loads / stores, SYNCS = mbarrier, DMMA = FP64 tensor-core MMA, DFMA/DMUL/DADD = FP64 pipe)
from `cuobjdump -sass` of quantumcomputer_b200/lib/libqcs.so.   python tools/sass_summary.py > profiles/<name>.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(anonymous namespace\)::", "", o) for o in out]


def main():
    rows = []
    for log in sorted(glob.glob(os.path.join(ROOT, "build", "obj", "*.ptxas.log"))):
        text = open(log).read()
        for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:ptxas info\s*:[^\n]*\n)*?"
                             r"\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                             r"ptxas info\s*: Used (\d+) registers(?:, used \d+ barriers)?(?:, \d+ bytes cumulative stack size)?(?:, (\d+) bytes smem)?", text):
            rows.append((os.path.basename(log).split(".")[0], m.group(1), int(m.group(5)), int(m.group(6) or 0),
                         int(m.group(2)), int(m.group(3)), int(m.group(4))))
    names = demangle([r[1] for r in rows])
    print("file            regs  static smem  stack  spill st/ld  kernel")
    for (f, _, regs, smem, stack, st, ld), name in zip(rows, names):
        name = re.sub(r"\(.*", "", name.replace("void ", ""))
        print(f"{f:15s} {regs:4d}  {smem:11d}  {stack:5d}  {st:5d}/{ld:<5d}  {name}")
    lib = os.path.join(ROOT, "quantumcomputer_b200", "lib", "libqcs.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    print("\narch:", sorted(set(re.findall(r"arch = (sm_\w+)", sass))))
    per = collections.defaultdict(collections.Counter)
    fn = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            per[fn][m.group(1)] += 1
    keys = ["UTMALDG", "UTMASTG", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "SHFL", "ATOMG", "RED"]
    fns = sorted(per)
    pretty = demangle(fns)
    print("\nSASS mnemonics per kernel (instructions in the binary, not executed counts)")
    print("  ".join(f"{k:>7s}" for k in keys) + "  kernel")
    for f, name in zip(fns, pretty):
        name = re.sub(r"\(.*", "", name.replace("void ", ""))
        print("  ".join(f"{per[f][k]:7d}" for k in keys) + "  " + name)
    total = collections.Counter()
    for f in fns:
        total.update(per[f])
    print("  ".join(f"{total[k]:7d}" for k in keys) + "  TOTAL")


if __name__ == "__main__":
    main()
