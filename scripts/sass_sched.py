"""Static issue-time model of one warp running the substep loop of orl_step_kernel (one warp per scheduler, so the
stall counts ptxas encoded ARE the issue times; /opt/skills/guides/B300_MICROARCH.md "single-warp issue model").

    python scripts/sass_sched.py <lib.so> [kernel-substring] [--blocks] [--trace]

Decodes the control word of every SASS instruction (stall, yield, write/read barrier, wait mask), finds the substep loop
(largest backward branch) and the rolled per-joint loop inside it, and reports per basic block: instructions, the sum of
the stall fields (= cycles if no scoreboard wait binds) and an event-model estimate with fixed latencies for the
variable-latency classes (LDS 29, MUFU 18, SHFL 24, LDG 300).  Rarely taken branches are assumed not taken.
"""
import re
import subprocess
import sys
import collections

LAT = {"LDS": 29, "MUFU": 18, "SHFL": 24, "LDG": 300, "LDC": 40, "LDCU": 40, "S2R": 30, "ATOM": 300, "RED": 10, "STS": 6, "STG": 6,
       "DFMA": 12, "DMUL": 12, "DADD": 12, "F2F": 12, "I2F": 12, "F2I": 12, "DSETP": 12, "LDL": 300, "STL": 6, "POPC": 12,
       "FLO": 12, "BREV": 12, "VOTE": 12, "REDUX": 30, "R2UR": 12, "LDSM": 30, "CS2R": 12, "S2UR": 30}


def parse(lib, kern):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(out) if "Function : " in l and kern in l][0]
    end = [i for i, l in enumerate(out) if "Function : " in l and i > start]
    end = end[0] if end else len(out)
    ins = []
    i = start
    while i < end:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", out[i])
        if m:
            m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", out[i + 1])
            hi = int(m2.group(1), 16)
            text = m.group(2).strip()
            op = text.split()[1] if text.startswith("@") else text.split()[0]
            ins.append(dict(addr=int(m.group(1), 16), text=text, op=op, base=op.split(".")[0], stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1,
                            wbar=(hi >> 46) & 7, rbar=(hi >> 49) & 7, wait=(hi >> 52) & 0x3F))
            i += 2
        else:
            i += 1
    return ins


def loops(ins):
    res = []
    for x in ins:
        m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", x["text"])
        if m and x["base"] == "BRA":
            tgt = int(m.group(1), 16)
            if tgt <= x["addr"]:
                res.append((x["addr"] - tgt + 16, tgt, x["addr"]))
    return sorted(res, reverse=True)


def simulate(seq):
    """Event model of one warp over a straight instruction sequence."""
    T = 0
    sb = [0] * 6
    waits = collections.Counter()
    for x in seq:
        arm = max([sb[s] for s in range(6) if (x["wait"] >> s) & 1] or [0])
        if arm > T:
            waits[x["base"]] += arm - T
            T = arm
        x["t"] = T
        lat = LAT.get(x["base"], 20)
        if x["wbar"] < 6:
            sb[x["wbar"]] = max(sb[x["wbar"]], T + lat)
        if x["rbar"] < 6:
            sb[x["rbar"]] = max(sb[x["rbar"]], T + 6)
        T += max(x["stall"], 1)
    return T, waits


def main():
    lib = sys.argv[1]
    kern = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "orl_step_kernelILb0ELb0E"
    ins = parse(lib, kern)
    by_addr = {x["addr"]: i for i, x in enumerate(ins)}
    lp = loops(ins)
    lp = [l for l in lp if l[0] // 16 <= 4000]      # the outermost backward branches are out-of-line divergence handlers
    nfp = lambda l: sum(1 for x in ins[by_addr[l[1]]:by_addr[l[2]] + 1] if x["base"] in ("FFMA", "FMUL", "FADD", "FFMA2"))
    lp.sort(key=nfp, reverse=True)                  # the substep loop is the one with the most FP32 instructions
    for a in sys.argv:
        if a.startswith("--loop="):                 # pick the loop whose head is at this address
            lp = [l for l in lp if l[1] == int(a[7:], 16)] + lp
    span, t0, t1 = lp[0]
    body = ins[by_addr[t0]:by_addr[t1] + 1]
    inner = [l for l in lp[1:] if l[1] >= t0 and l[2] <= t1 and (l[1], l[2]) != (t0, t1)]
    print("kernel %s: %d instructions; substep loop 0x%x..0x%x = %d instructions (%d B)" % (kern, len(ins), t0, t1, len(body), span))
    for l in inner:
        print("  inner loop 0x%x..0x%x = %d instructions" % (l[1], l[2], l[0] // 16))
    # basic blocks: leaders = branch targets and instructions after branches
    leaders = {t0}
    for x in body:
        m = re.search(r"0x([0-9a-f]+)", x["text"]) if x["base"] in ("BRA", "BSSY", "CALL") else None
        if x["base"] in ("BRA", "EXIT", "RET", "BRX"):
            leaders.add(x["addr"] + 16)
        if m and x["base"] == "BRA":
            leaders.add(int(m.group(1), 16))
    blocks = []
    cur = []
    for x in body:
        if x["addr"] in leaders and cur:
            blocks.append(cur); cur = []
        cur.append(x)
    if cur:
        blocks.append(cur)
    if "--blocks" in sys.argv:
        for b in blocks:
            T, w = simulate([dict(y) for y in b])
            fp = sum(1 for y in b if y["base"] in ("FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2"))
            print("  block 0x%05x n=%4d fp=%4d stall-sum=%5d model=%5d  last: %s" % (b[0]["addr"], len(b), fp, sum(max(y["stall"], 1) for y in b), T, b[-1]["text"][:60]))
    # fall-through path: take the per-joint loop 3 times, other backward branches once, skip predicated forward branches
    main_inner = max(inner, key=lambda l: l[0]) if inner else None
    seq = []
    i = by_addr[t0]
    trips = 0
    guard = 0
    while i <= by_addr[t1] and guard < 100000:
        guard += 1
        x = ins[i]
        seq.append(dict(x))
        if x["base"] == "BRA" and re.search(r"0x([0-9a-f]+)", x["text"]):
            m = re.search(r"0x([0-9a-f]+)", x["text"])
            tgt = int(m.group(1), 16)
            pred = x["text"].startswith("@")
            if main_inner and x["addr"] == main_inner[2]:
                trips += 1
                if trips < 3:
                    i = by_addr[tgt]; continue
            elif not pred and tgt > x["addr"]:
                i = by_addr[tgt]; continue           # unconditional forward branch
            elif pred and tgt > x["addr"] and any(l[1] > x["addr"] and l[2] < tgt for l in inner):
                i = by_addr[tgt]; continue           # a predicated forward branch around a rolled (rare) loop: taken
        i += 1
    T, w = simulate(seq)
    n = len(seq)
    cls = collections.Counter(x["base"] for x in seq)
    print("fall-through path of one substep: %d instructions, stall-sum %d, event model %d cycles (%.2f cycles/instr)" % (
        n, sum(max(x["stall"], 1) for x in seq), T, T / n))
    print("  scoreboard waits by consumer class:", dict(w.most_common(8)))
    print("  instruction mix:", dict(cls.most_common(14)))
    hist = collections.Counter(max(x["stall"], 1) for x in seq)
    print("  stall-field histogram:", dict(sorted(hist.items())))
    fp = [x for x in seq if x["base"] in ("FFMA", "FMUL", "FADD")]
    print("  FP32 instructions %d; their stall-sum %d" % (len(fp), sum(max(x["stall"], 1) for x in fp)))
    if "--trace" in sys.argv:
        for x in seq:
            print("%6d %05x s=%2d w=%02x wb=%d rb=%d  %s" % (x["t"], x["addr"], x["stall"], x["wait"], x["wbar"], x["rbar"], x["text"][:90]))


if __name__ == "__main__":
    main()
