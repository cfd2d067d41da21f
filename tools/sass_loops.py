#!/usr/bin/env python
"""Static view of a kernel's SASS: opcode histogram of the whole kernel and of every loop (backward branch).
usage: tools/sass_loops.py [lib.so|obj.o] [kernel-substring] [min-loop-size]
ALU = the 16-lane integer pipe (VIADDMNMX, VIMNMX*, ISETP, SEL, LOP3, IADD3, VIADD, SHF, LEA, PRMT ...),
FMA = IMAD* (the other 16-lane pipe); measured rates: profiles/r02_pipe_rates.txt."""
import collections, os, re, subprocess, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "centrolign_b200", "csrc", "libcentrolign_b200.so")
kern = sys.argv[2] if len(sys.argv) > 2 else "popoa_kernelILi3"
minsz = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dump = sys.argv[4] if len(sys.argv) > 4 else None  # "0xBEGIN-0xEND": print that address range instead

ALU = {"VIADDMNMX", "VIMNMX", "VIMNMX3", "ISETP", "SEL", "LOP3", "IADD3", "VIADD", "SHF", "LEA", "PRMT", "PLOP3", "IABS", "FMNMX", "MOV", "P2R", "R2P", "POPC", "FLO", "BREV", "IMNMX"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD"}
MEM = {"LDS", "STS", "LDG", "STG", "LD", "ST", "LDGSTS", "LDSM", "ATOMS", "ATOMG", "RED", "LDC", "LDCU", "LDL", "STL"}
CTL = {"BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "BAR", "NANOSLEEP", "YIELD", "BREAK", "NOP"}


def cls(op):
    b = op.split(".")[0]
    if b in ALU: return "alu"
    if b in FMA: return "fma"
    if b in MEM: return "mem"
    if b in CTL: return "ctl"
    if b == "SHFL": return "shfl"
    return "oth"


out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
fn, ins = None, []
for l in out.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1)
        continue
    if fn and kern in fn:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            addr, text = int(m.group(1), 16), m.group(2).strip()
            mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", text)
            ins.append((addr, mm.group(2) if mm else text, text))
if dump:
    lo, hi = (int(x, 16) for x in dump.split("-"))
    for a, o, t in ins:
        if lo <= a <= hi:
            print(f"{a:#07x}  {t}")
    sys.exit(0)
print(f"{kern}: {len(ins)} instructions")
idx = {a: i for i, (a, _, _) in enumerate(ins)}
tot = collections.Counter(cls(o) for _, o, _ in ins)
print("whole kernel:", dict(tot))
loops = []
for i, (a, op, text) in enumerate(ins):
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", text)
        if m:
            t = int(m.group(1), 16)
            if t in idx and idx[t] <= i:
                loops.append((idx[t], i))
for (b, e) in sorted(set(loops)):
    n = e - b + 1
    if n < minsz:
        continue
    c = collections.Counter(cls(o) for _, o, _ in ins[b:e + 1])
    ops = collections.Counter(o.split(".")[0] for _, o, _ in ins[b:e + 1])
    print(f"loop {ins[b][0]:#07x}..{ins[e][0]:#07x}: {n} instr  {dict(c)}")
    print("    ", ", ".join(f"{k} {v}" for k, v in ops.most_common(18)))
