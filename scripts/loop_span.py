"""Static size of orl_step_kernel and of its substep loop (largest backward branch span) in a built library."""
import re, subprocess, sys
for lib in sys.argv[1:]:
    lines = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(lines) if "Function : " in l and "orl_step_kernel" in l][0]
    end = [i for i, l in enumerate(lines) if "Function : " in l and i > start]
    end = end[0] if end else len(lines)
    ins = []
    for l in lines[start:end]:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    best = None
    for a, t in ins:
        m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[0]):
                best = (a - tgt, tgt, a)
    print("%-36s kernel %6d B  substep loop span %6d B = %5d instr" % (lib.split("/")[-1], ins[-1][0] + 16, best[0] + 16, (best[0] + 16) // 16))
