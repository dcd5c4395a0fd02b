"""Development aid: attribute an ncu source-page capture of trace_kernel to the PHASES of a pass.

    ncu -i X.ncu-rep --page source --csv --print-source sass > X.sass.csv
    python scripts/ncu_by_phase.py X.sass.csv [--cubin K.cubin | --flags "-DRT_..."] [--kernel REGEX]

Every SASS instruction carries its inlining chain (nvdisasm -gi, needs -lineinfo).  The chain is
walked from the kernel body inwards; the first frame that lies in a pass / advance / helper function
names the phase, and inside pass_* the line decides between set-up, filter loop, gather and resolve.
Out-of-line helpers (IEEE division / square root, the exact test) are reported under their own name.
Prints, per phase: share of executed warp instructions, share of stall samples, lane efficiency
(thread instructions / 32 x warp instructions)."""
import argparse, bisect, collections, csv, re, subprocess, sys, tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "raytracer-gamma_b200" / "csrc"
ap = argparse.ArgumentParser()
ap.add_argument("sass_csv")
ap.add_argument("--cubin")
ap.add_argument("--flags", default="")
ap.add_argument("--kernel", default=None, help="regex on the mangled kernel name (default: taken from the capture)")
ap.add_argument("--lines", type=int, default=0, help="also print the N hottest source lines")
ap.add_argument("--stalls", action="store_true", help="per phase: the stall reasons of its samples")
ap.add_argument("--root", default=None, help="source tree of the captured build (default: this checkout)")
args = ap.parse_args()
if args.root:
    ROOT = Path(args.root).resolve()
    CSRC = ROOT / "raytracer-gamma_b200" / "csrc"

STATIC = args.sass_csv == "-"        # no capture: static instruction counts per phase only (needs --kernel)
if STATIC:
    kname, hdr, data, ix = args.kernel, [], None, {}
else:
    rows = list(csv.reader(open(args.sass_csv)))
    kname = rows[0][1]
    hdr, data = rows[1], rows[2:]
    ix = {n: i for i, n in enumerate(hdr)}

cubin = args.cubin
if not cubin:
    tmp = Path(tempfile.mkdtemp())
    cubin = str(tmp / "k.cubin")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
                    *args.flags.split(), f"-I{ROOT/'include'}", f"-I{CSRC}", "-cubin", "-o", cubin, str(CSRC / "rt_shim.cu")],
                   check=True, stderr=subprocess.DEVNULL)
dis = subprocess.run(["nvdisasm", "-gi", cubin], capture_output=True, text=True).stdout

# ---- which text section is the captured kernel?  match template arguments of the demangled name
def demangle(m):
    return subprocess.run(["cu++filt", m], capture_output=True, text=True).stdout.strip()
sections = re.findall(r"^\.text\.(\S+):", dis, flags=re.M)
want = re.sub(r"\s+", "", kname).replace("(bool)", "").replace("(int)", "")
target = None
for s in sections:
    d = re.sub(r"\s+", "", demangle(s)).replace("(bool)", "").replace("(int)", "")
    if args.kernel:
        if re.search(args.kernel, s):
            target = s; break
    elif d == want:
        target = s; break
if target is None:
    sys.exit(f"kernel {kname} not found in the cubin (sections: {[demangle(s) for s in sections if 'trace' in s]})")

# ---- per instruction: inlining chain and enclosing sub-function label
insts = []   # (frames innermost->outermost as (file, line), subfunc or None)
active = False; frames = []; pending = []; sub = None
for line in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        active = (m.group(1) == target); sub = None; continue
    if not active:
        continue
    m = re.match(r"\$" + re.escape(target) + r"\$(\S+):", line.strip())
    if m:
        sub = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at)?', line)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        if not m.group(3):
            frames = pending; pending = []
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        insts.append((frames, sub))
if STATIC:
    data = [None] * len(insts)
assert len(insts) == len(data), (len(insts), len(data), "the cubin is not the captured build")

def funcs_of(path):
    out = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"(?:RT_HD(?:_NI)?|__device__ __forceinline__|__device__ __noinline__|__global__)\s+.*?(\w+)\(", l) or \
            re.match(r"(?:template.*>\s*)?__device__.*?\s(\w+)\(", l)
        if m:
            out.append((i, m.group(1)))
    return out
SRC = {f: open(CSRC / f).read().split("\n") for f in ("rt_core.cuh", "rt_kernels.cuh")}
FN = {f: funcs_of(CSRC / f) for f in SRC}
def fn_at(f, l):
    if f not in FN:
        return None
    j = bisect.bisect_right([x[0] for x in FN[f]], l) - 1
    return FN[f][j][1] if j >= 0 else None
KERNEL_FUNCS = {"trace_body", "trace_kernel", "trace_kernel_const", "__launch_bounds__"}

def lambda_kind(f, l):
    """a line inside one of the pass lambdas: which one?  (looks backwards for the lambda header)"""
    src = SRC[f]
    for k in range(l - 1, max(0, l - 120), -1):
        t = src[k]
        if "[&](uint32_t base)" in t: return "filter loop"
        if "[&](uint32_t e)" in t: return "resolve"
        if "[&](uint32_t cl" in t: return "members"
        if re.search(r"filter_rounds<|accel_rounds<", t): return None
        if re.search(r"__device__ __forceinline__ void pass_", t): return None
    return None

def rounds_phase(f, l):
    """a line of filter_rounds / accel_rounds itself"""
    src = SRC[f]
    for k in range(l - 1, max(0, l - 80), -1):
        t = src[k]
        if re.search(r"for \(int (k|j) = 0", t) or "resolve(" in t: return "resolve"
        if "members(" in t: return "members"
        if re.search(r"for \((uint32_t base|; base < nPad)", t): return "filter loop"
        if "__device__ __forceinline__ void" in t: return "filter loop"
    return "filter loop"

def kernel_phase(l):
    src = SRC["rt_kernels.cuh"]
    for k in range(l - 1, max(0, l - 400), -1):
        t = src[k]
        if "---- per-warp reductions" in t: return "kernel: epilogue"
        if "---- advance the served slots" in t: return "kernel: advance glue"
        if "---- vote ----" in t: return "kernel: vote + dispatch"
        if "---- refill" in t: return "kernel: refill"
        if "void trace_body(" in t: return "kernel: prologue"
    return "kernel: other"

ADVANCE = ("advance_slot", "advance", "after_matte", "after_contain", "unwind", "fresnel_term", "setup_shadow_batch",
           "load_slot", "store_slot", "sample_value")

def classify(frames, sub):
    if sub:
        d = demangle(sub)
        return "out-of-line: " + re.sub(r"\(.*", "", d).replace("rtg::", "")
    chain = list(reversed(frames))          # outermost first
    fns = [(f, l, fn_at(f, l)) for f, l in chain]
    names = [fn for _, _, fn in fns]
    if any(fn in ("gather", "gather1", "gather2") for fn in names):
        return "gather"
    for f, l, fn in fns:
        if fn is None or fn in KERNEL_FUNCS:
            continue
        if fn.startswith("pass_"):
            # the deepest frame that still lies in this pass function decides (lambda bodies are lines of pass_*)
            deep = [(ff, ll) for ff, ll, nn in fns if nn == fn]
            kind = lambda_kind(*deep[-1])
            if kind is None and any(nn in ("filter_rounds", "accel_rounds") for nn in names):
                fr = [(ff, ll) for ff, ll, nn in fns if nn in ("filter_rounds", "accel_rounds")][-1]
                kind = rounds_phase(*fr)
            if kind is None:
                kind = "set-up / epilogue"
            return f"{fn}: {kind}"
        if fn in ADVANCE:
            return "advance (state machine)"
        if fn in ("start_task", "work_to_task", "set_trace_query"):
            return "kernel: refill"
        return "other: " + fn
    inner = [(f, l) for f, l, fn in fns if fn == "trace_body"]
    return kernel_phase(inner[-1][1]) if inner else "kernel: other"

STALLS = ["stall_no_inst", "stall_long_sb", "stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_branch_resolving",
          "stall_short_sb", "stall_dispatch", "stall_lg", "stall_mio", "stall_barrier"]
stall = collections.defaultdict(collections.Counter)
inst = collections.Counter(); samp = collections.Counter(); thr = collections.Counter(); byline = collections.Counter(); static = collections.Counter()
ti = ts = 0
for r, (frames, sub) in zip(data, insts):
    a, b, t = (1, 1, 32) if STATIC else (int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), int(r[ix["Thread Instructions Executed"]]))
    key = classify(frames, sub)
    if not STATIC and args.stalls:
        for sname in STALLS:
            if sname in ix and r[ix[sname]]:
                stall[key][sname] += int(r[ix[sname]])
    inst[key] += a; samp[key] += b; thr[key] += t; ti += a; ts += b; static[key] += 1
    if frames:
        byline[frames[-1] if frames[-1][0] == "rt_kernels.cuh" else frames[0]] += b
print(f"kernel {kname}")
print(f"warp instructions {ti}  stall samples {ts}")
print(f"{'phase':44s} {'inst %':>8s} {'samples %':>10s} {'lane eff':>9s} {'SASS':>6s}")
for k, c in sorted(samp.items(), key=lambda kv: -kv[1]):
    if inst[k] == 0 and c == 0:
        continue
    print(f"{k:44s} {100 * inst[k] / ti:8.2f} {100 * c / ts:10.2f} {thr[k] / max(1, 32 * inst[k]):9.3f} {static[k]:6d}")
if args.stalls and not STATIC:
    print("stall reasons, % of ALL samples of the kernel (rows: phases; columns: " + " ".join(x[6:] for x in STALLS) + ")")
    tot = collections.Counter()
    for k, c in sorted(samp.items(), key=lambda kv: -kv[1]):
        if c * 200 < ts:
            continue
        print(f"{k:44s} " + " ".join(f"{100 * stall[k][x] / ts:6.2f}" for x in STALLS))
    for k in stall:
        for x in STALLS:
            tot[x] += stall[k][x]
    print(f"{'TOTAL':44s} " + " ".join(f"{100 * tot[x] / ts:6.2f}" for x in STALLS))
print(f"static SASS instructions {len(data)} ({len(data) * 16 / 1024:.1f} KB)")
grp = collections.Counter(); grs = collections.Counter()
for k in inst:
    g = "filter loops" if k.endswith("filter loop") else "everything else"
    grp[g] += inst[k]; grs[g] += samp[k]
for g in grp:
    print(f"== {g:40s} {100 * grp[g] / ti:8.2f} {100 * grs[g] / ts:10.2f}")
if args.lines:
    print("hottest lines (stall samples, outermost frame in rt_kernels.cuh)")
    for (f, l), c in byline.most_common(args.lines):
        print(f"{100 * c / ts:6.2f}%  {f}:{l}  {SRC.get(f, [''] * (l + 1))[l - 1].strip()[:90]}")
