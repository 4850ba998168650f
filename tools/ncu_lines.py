"""Top source lines of one kernel in an .ncu-rep by stall samples / executed instructions.
usage: ncu_lines.py report.ncu-rep kernel_name object.o [top]   (object.o = the -lineinfo object the kernel came from)"""
import csv, io, os, re, subprocess, sys, tempfile
rep, kern, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# ---- SASS offset -> (file, line) from nvdisasm
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = kern in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
# ---- per-SASS metrics from the report
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = next(r for r in rows if r and r[0] == "Address")
agg, base = {}, None
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows:
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    d = dict(zip(hdr, r))
    addr = int(d["Address"], 16)
    base = addr if base is None else base
    key = line_of.get(addr - base, ("?", 0))
    a = agg.setdefault(key, {"s": 0, "i": 0, "st": {}})
    a["s"] += int(d["# Samples"]); a["i"] += int(d["Instructions Executed"])
    for c in stall_cols:
        v = int(d[c] or 0)
        if v:
            a["st"][c] = a["st"].get(c, 0) + v
tot = sum(a["s"] for a in agg.values()) or 1
toti = sum(a["i"] for a in agg.values()) or 1
src = {}
print("total samples", tot, "warp instructions", toti)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:top]:
    f, l = key
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:80] if 0 < l <= len(src[f]) else ""
    st = ",".join("%s:%d" % (k.replace("stall_", ""), v) for k, v in sorted(a["st"].items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% smp %5.1f%% inst %s:%d  %s   [%s]" % (100 * a["s"] / tot, 100 * a["i"] / toti, f, l, text, st))
