"""Per-region summary of an ncu source-page export of orl_step_kernel:
    ncu -i rep.ncu-rep --page source --csv > source.csv ; python scripts/ncu_regions.py source.csv <lib.so>
Regions are address ranges of the SASS (substep loop found by the same rule as sass_sched.py); per region: samples, share,
executed warp instructions, samples per instruction and the stall-reason mix.  `--lines` adds the hottest source lines
outside the loop (needs a nvdisasm -gi listing passed as third argument)."""
import csv, sys, collections, re, subprocess
sys.path.insert(0, __file__.rsplit("/", 1)[0])
import sass_sched

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
base = int(data[0][col["Address"]], 16)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
ins = sass_sched.parse(sys.argv[2], "orl_step_kernelILb0ELb0E")
by_addr = {x["addr"]: i for i, x in enumerate(ins)}
lp = [l for l in sass_sched.loops(ins) if l[0] // 16 <= 4000]
nfp = lambda l: sum(1 for x in ins[by_addr[l[1]]:by_addr[l[2]] + 1] if x["base"] in ("FFMA", "FMUL", "FADD"))
lp.sort(key=nfp, reverse=True)
_, t0, t1 = lp[0]
regions = collections.OrderedDict()
def region(a):
    if a < t0: return "before loop"
    if a <= t1: return "substep loop"
    return "after loop"
agg = collections.defaultdict(lambda: collections.Counter())
tot = 0
per_addr = {}
for r in data:
    a = int(r[col["Address"]], 16) - base
    n = int(r[col["# Samples"]] or 0)
    ex = int(r[col["Instructions Executed"]] or 0)
    g = region(a)
    agg[g]["samples"] += n; agg[g]["executed"] += ex; agg[g]["static"] += 1
    for s in stalls:
        agg[g][s] += int(r[col[s]] or 0)
    tot += n
    per_addr[a] = (n, ex, r[col["Source"]].strip(), {s: int(r[col[s]] or 0) for s in stalls})
print("total samples", tot)
for g, c in agg.items():
    mix = sorted(((c[s], s[6:]) for s in stalls), reverse=True)[:6]
    print("%-13s samples %6d (%4.1f %%)  static %5d  executed %9d  samples/1k-exec %.2f  | %s" % (
        g, c["samples"], 100.0 * c["samples"] / tot, c["static"], c["executed"], 1000.0 * c["samples"] / max(c["executed"], 1),
        "  ".join("%s %.0f%%" % (s, 100.0 * v / max(c["samples"], 1)) for v, s in mix)))
if "--top" in sys.argv:
    # hottest 64-instruction windows after the loop
    addrs = sorted(a for a in per_addr if a > t1)
    win = 64
    out = []
    for i in range(0, len(addrs), win):
        w = addrs[i:i + win]
        n = sum(per_addr[a][0] for a in w); ex = sum(per_addr[a][1] for a in w)
        st = collections.Counter()
        for a in w:
            st.update(per_addr[a][3])
        out.append((n, w[0], ex, st))
    for n, a0, ex, st in sorted(out, reverse=True)[:30]:
        print("0x%05x  samples %5d  executed %7d  %s" % (a0, n, ex, " ".join("%s:%d" % (k[6:], v) for k, v in st.most_common(4))))
