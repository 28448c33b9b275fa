"""Per-source-line view of an ncu report (the --page source CSV is SASS-level): joins it with nvdisasm's line table of the
same object by instruction order.  usage: ncu_lines.py <report.ncu-rep> <object.o> <mangled-kernel-substring> [top] [demangled-substring]"""
import csv, re, subprocess, sys, tempfile, os, glob
from collections import defaultdict

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
lines = []  # (file, line) per instruction of the kernel, in order
for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
    out = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout
    cur, fn = None, None
    for ln in out.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            fn = m.group(1)
            cur = None
            continue
        if fn and kname in fn and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
            lines.append((cur, ln.split("*/", 1)[1].strip().rstrip(";")))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
# several kernels may be in the report: take the first block whose name matches
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
human = sys.argv[5] if len(sys.argv) > 5 else kname  # substring of the demangled name as ncu prints it
blk = [b for b in blocks if human in b["name"]][0]
hdr = blk["hdr"]
ix = {h: i for i, h in enumerate(hdr)}
n = min(len(lines), len(blk["rows"]))
if len(lines) != len(blk["rows"]):
    print(f"warning: {len(lines)} SASS instructions in the object, {len(blk['rows'])} in the report", file=sys.stderr)
agg = defaultdict(lambda: defaultdict(float))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for (src, sass), r in zip(lines[:n], blk["rows"][:n]):
    a = agg[src]
    a["inst"] += float(r[ix["Instructions Executed"]] or 0)
    a["tinst"] += float(r[ix["Thread Instructions Executed"]] or 0)
    a["samp"] += float(r[ix["# Samples"]] or 0)
    for s in stall_cols:
        a[s] += float(r[ix[s]] or 0)
tot_i = sum(a["inst"] for a in agg.values()); tot_s = sum(a["samp"] for a in agg.values())
print(f"total inst {tot_i:.0f} samples {tot_s:.0f}")
srcs = {}
def text(src):
    if src is None: return ""
    f, l = src
    if f not in srcs:
        p = [x for x in glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "**", f), recursive=True)]
        srcs[f] = open(p[0]).read().splitlines() if p else []
    return srcs[f][l - 1].strip()[:90] if 0 < l <= len(srcs[f]) else ""
for src, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = sorted(((a[s], s[6:]) for s in stall_cols), reverse=True)[:3]
    print(f"{(src[0] if src else '?'):16s} {(src[1] if src else 0):5d} inst {a['inst']:12.0f} ({100*a['inst']/tot_i:4.1f}%) thr/inst {a['tinst']/max(1,a['inst']):4.1f} samp {a['samp']:7.0f} ({100*a['samp']/tot_s:4.1f}%) "
          + " ".join(f"{n}:{v:.0f}" for v, n in st if v > 0) + "  | " + text(src))
