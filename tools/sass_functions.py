"""Split the SASS of one kernel (cuobjdump -sass -fun <kernel>) into the out-of-line functions it calls and print
their size, FP64 instruction count, and callers.  usage: python tools/sass_functions.py kernel.sass"""
import re, sys, collections
ins = []  # (addr, text)
for line in open(sys.argv[1]):
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
targets = collections.defaultdict(list)
for a, t in ins:
    m = re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", t)
    if m:
        targets[int(m.group(1), 16)].append(a)
starts = sorted(targets)
bounds = [0] + starts + [ins[-1][0] + 16]
def fn_of(addr):
    k = 0
    for i, s in enumerate(bounds[:-1]):
        if addr >= s: k = i
    return k
names = {0: "kernel body"}
for i, s in enumerate(starts): names[i + 1] = f"fn@{s:#x}"
stat = collections.defaultdict(lambda: collections.Counter())
for a, t in ins:
    k = fn_of(a)
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    stat[k]["n"] += 1
    if re.match(r"D(FMA|ADD|MUL|SETP|MNMX)", op): stat[k]["fp64"] += 1
    if op.startswith("MUFU"): stat[k]["mufu"] += 1
    if op.startswith(("LDS", "STS")): stat[k]["lds"] += 1
    if op.startswith(("LDG", "LD.", "LDL", "STL")): stat[k]["ldg/l"] += 1
    if op.startswith("NOP"): stat[k]["nop"] += 1
print("| function | instr | fp64 | mufu | lds/sts | ldg/ldl/stl | nop | called from |")
for k in sorted(stat):
    c = stat[k]
    callers = collections.Counter(names[fn_of(a)] for a in targets.get(bounds[k], [])) if k else {}
    print(f"| {names[k]} | {c['n']} | {c['fp64']} | {c['mufu']} | {c['lds']} | {c['ldg/l']} | {c['nop']} | {dict(callers)} |")
