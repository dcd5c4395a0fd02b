"""Development aid: join an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv --print-source sass)
with nvdisasm line info of the same build, and sum executed instructions / stall samples per source function.
usage: ncu_by_function.py sass.csv mangled_kernel_name"""
import bisect, collections, csv, re, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
src_csv, target = sys.argv[1], sys.argv[2]
tmp = Path(tempfile.mkdtemp())
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
                f"-I{ROOT/'include'}", f"-I{ROOT/'raytracer-gamma_b200/csrc'}", "-cubin", "-o", str(tmp/"k.cubin"),
                str(ROOT/"raytracer-gamma_b200/csrc/rt_shim.cu")], check=True, stderr=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "--print-line-info", str(tmp/"k.cubin")], capture_output=True, text=True).stdout
lines = []; active = False; cur = None
for line in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        active = (m.group(1) == target); continue
    if not active: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line): lines.append(cur)
rows = list(csv.reader(open(src_csv)))
h = rows[1]; data = rows[2:]; ix = {n: i for i, n in enumerate(h)}
assert len(data) == len(lines), (len(data), len(lines))
def funcs_of(path):
    out = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"(?:RT_HD(?:_NI)?|__device__ __forceinline__|__device__ __noinline__|__global__)\s+.*?(\w+)\(", l) or \
            re.match(r"(?:template.*>\s*)?__device__.*?\s(\w+)\(", l)
        if m: out.append((i, m.group(1)))
    return out
F = {f: funcs_of(ROOT/"raytracer-gamma_b200/csrc"/f) for f in ("rt_core.cuh", "rt_kernels.cuh")}
inst = collections.Counter(); samp = collections.Counter(); byline = collections.Counter()
ti = ts = 0
for r, loc in zip(data, lines):
    a = int(r[ix["Instructions Executed"]]); b = int(r[ix["# Samples"]])
    ti += a; ts += b
    key = "?"
    if loc and loc[0] in F:
        fs = F[loc[0]]; j = bisect.bisect_right([f[0] for f in fs], loc[1]) - 1
        key = f"{loc[0]}:{fs[j][1] if j >= 0 else '?'}"
        byline[(loc[0], loc[1])] += b
    elif loc: key = loc[0]
    inst[key] += a; samp[key] += b
print("total inst", ti, "samples", ts)
for k, c in samp.most_common(30): print(f"{100*c/ts:6.2f}% samples {100*inst[k]/ti:6.2f}% inst  {k}")
print("hottest lines (samples)")
for (f, l), c in byline.most_common(25): print(f"{100*c/ts:6.2f}%  {f}:{l}")
