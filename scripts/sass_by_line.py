"""Static SASS instruction count of orl_step_kernel attributed to the OUTERMOST source line in orl_kernels.cuh
(walks nvdisasm -gi inline chains).  Usage: sass_by_line.py <disassembly from `nvdisasm -gi cubin`> [kernel]"""
import re, sys, collections
lines = open(sys.argv[1]).read().split("\n")
kern = sys.argv[2] if len(sys.argv) > 2 else "orl_step_kernel"
SKIP = {int(a.split("=")[1]) for a in sys.argv if a.startswith("--skip-line=")}     # e.g. the kernel wrapper's call of step_worker
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
end = [i for i, l in enumerate(lines) if l.startswith(".text.") and i > start]
end = end[0] if end else len(lines)
cur = None
chain = []
cnt = collections.Counter()
total = 0
for l in lines[start:end]:
    if "//##" in l:
        # consecutive //## lines form one inline chain, innermost location first; keep every location of the chain
        chain.extend((f.split("/")[-1], int(n)) for f, n in re.findall(r'File "([^"]+)", line (\d+)', l))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        if chain:
            outer = [c for c in chain if c[0] == "orl_kernels.cuh" and c[1] not in SKIP]
            cur = outer[-1] if outer else chain[-1]
            chain = []
        cnt[cur] += 1
        total += 1
print("total", total)
for k in sorted(cnt, key=lambda k: (k[0], k[1])):
    if cnt[k] >= 15:
        print("%5d  %s:%d" % (cnt[k], k[0], k[1]))
