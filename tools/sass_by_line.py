"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report.

ncu's `--page source --csv` lists the SASS of a kernel with its per-instruction counters but without
source lines; `nvdisasm -g` on the cubin inside libpdmops.so (built with -lineinfo) lists the same
SASS with `//## File ..., line N` markers.  The two listings are joined by instruction order.

    python tools/sass_by_line.py gpurun_out/x.ncu-rep <kernel regex> [--so pdm_ssd_b200/libpdmops.so] [--top 40]
"""
import argparse, collections, csv, glob, io, os, re, subprocess, sys, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("kernel")
ap.add_argument("--so", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pdm_ssd_b200", "libpdmops.so"))
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", "regex:" + a.kernel],
                     capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
name = rows[0][1]
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # only the first captured launch
    if len(r) == len(hdr) and r[0] != "Address":
        data.append(dict(zip(hdr, r)))
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", name.split("(")[0].split("::")[-1].split("<")[0])

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.so)], cwd=tmp, check=True, capture_output=True)
best = None
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    cur, ins, fn = None, [], None
    funcs = {}
    for l in txt.split("\n"):
        m = re.match(r"\.text\.(\S+):", l)
        if m:
            fn = m.group(1)
            funcs[fn] = []
            cur = None
            continue
        m = re.search(r'//## File "(.*)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and fn:
            funcs[fn].append((m.group(2).strip(), cur))
    for fn, ins in funcs.items():
        if mangled_hint in fn and len(ins) == len(data):
            mism = 0
            for (t, _), d in zip(ins, data):
                o1 = t.split()[1] if t.startswith("@") else t.split()[0]
                s = d["Source"].strip()
                o2 = s.split()[1] if s.startswith("@") else s.split()[0]
                mism += o1.split(".")[0] != o2.split(".")[0]
            if best is None or mism < best[0]:
                best = (mism, fn, ins)
if best is None:
    sys.exit("no function in %s matches %s with %d instructions (rebuild the .so the report was taken with)" % (a.so, mangled_hint, len(data)))
mism, fn, ins = best


def gi(d, k):
    try:
        return int(d[k])
    except (ValueError, KeyError):
        return 0


byline, smp = collections.Counter(), collections.Counter()
for (t, ln), d in zip(ins, data):
    byline[ln] += gi(d, "Instructions Executed")
    smp[ln] += gi(d, "# Samples")
tot, ts = sum(byline.values()), max(1, sum(smp.values()))
src_cache = {}


def src_line(ln):
    if ln is None:
        return ""
    f, n = ln
    if f not in src_cache:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(a.so)), "csrc", f))
        src_cache[f] = open(cands[0]).read().split("\n") if cands else None
    return src_cache[f][n - 1].strip()[:100] if src_cache[f] and n <= len(src_cache[f]) else ""


print("# %s" % name[:150])
print("# %d SASS instructions, %d warp instructions executed, %d stall samples; opcode mismatches in the join: %d" % (len(data), tot, ts, mism))
print("# instr%  samples%  file:line  source")
for ln, c in sorted(byline.items(), key=lambda x: -x[1])[: a.top]:
    print("%6.2f  %6.2f  %s:%s  %s" % (100.0 * c / tot, 100.0 * smp[ln] / ts, ln[0] if ln else "?", ln[1] if ln else "?", src_line(ln)))
