"""Development aid: attribute the SASS instructions of one kernel to source functions (needs -lineinfo)."""
import bisect, collections, re, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
target = sys.argv[1] if len(sys.argv) > 1 else "_ZN3rtg12trace_kernelILb0ELi2ELi3EEEvNS_11TraceParamsE"
tmp = Path(tempfile.mkdtemp())
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false",
                f"-I{ROOT/'include'}", f"-I{ROOT/'raytracer-gamma_b200/csrc'}", "-cubin", "-o", str(tmp/"k.cubin"),
                str(ROOT/"raytracer-gamma_b200/csrc/rt_shim.cu")], check=True, stderr=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "--print-line-info", str(tmp/"k.cubin")], capture_output=True, text=True).stdout
cnt = collections.Counter(); active = False; cur = None; total = 0
for line in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        active = (m.group(1) == target); continue
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line) and cur:
        cnt[cur] += 1; total += 1
def funcs_of(path):
    out = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"(?:RT_HD(?:_NI)?|__device__ __forceinline__|__device__ __noinline__|__global__)\s+.*?(\w+)\(", l) or \
            re.match(r"(?:template.*>\s*)?__device__.*?\s(\w+)\(", l)
        if m: out.append((i, m.group(1)))
    return out
byfn = collections.Counter()
for fname in ("rt_core.cuh", "rt_kernels.cuh"):
    fs = funcs_of(ROOT/"raytracer-gamma_b200/csrc"/fname); starts = [f[0] for f in fs]
    for (f, l), c in cnt.items():
        if f == fname:
            j = bisect.bisect_right(starts, l) - 1
            byfn[(fname, fs[j][1] if j >= 0 else "?")] += c
other = sum(c for (f, l), c in cnt.items() if f not in ("rt_core.cuh", "rt_kernels.cuh"))
print("total", total, "other-headers", other)
for k, c in byfn.most_common(40): print(f"{c:6d} {k[0]}:{k[1]}")
