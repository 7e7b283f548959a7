"""Static SASS instruction count of orl_step_kernel attributed to the OUTERMOST source line in orl_kernels.cuh
(walks nvdisasm -gi inline chains).  Usage: sass_by_line.py <disassembly from `nvdisasm -gi cubin`> [kernel]"""
import re, sys, collections
lines = open(sys.argv[1]).read().split("\n")
kern = sys.argv[2] if len(sys.argv) > 2 else "orl_step_kernel"
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
end = [i for i, l in enumerate(lines) if l.startswith(".text.") and i > start]
end = end[0] if end else len(lines)
cur = None
cnt = collections.Counter()
total = 0
for l in lines[start:end]:
    if "//##" in l:
        # the chain lists the innermost location first, "inlined at" the callers after it
        locs = re.findall(r'File "([^"]+)", line (\d+)', l)
        outer = [(f.split("/")[-1], int(n)) for f, n in locs if f.endswith("orl_kernels.cuh")]
        cur = outer[-1] if outer else (locs[-1][0].split("/")[-1], int(locs[-1][1]))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        cnt[cur] += 1
        total += 1
print("total", total)
for k in sorted(cnt, key=lambda k: (k[0], k[1])):
    if cnt[k] >= 15:
        print("%5d  %s:%d" % (cnt[k], k[0], k[1]))
