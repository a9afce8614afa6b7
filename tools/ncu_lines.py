"""Attribute ncu warp-stall samples of one kernel to CUDA source lines (joins `ncu --page source` SASS rows with
`nvdisasm -g` line info of the in-tree library).  usage: python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <mangled substr> <src.cu>"""
import collections, csv, re, subprocess, sys, tempfile, os, glob
rep, kre, mangled, srcfile = sys.argv[1:5]
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all /root/repo/tensorrtx_b200/lib/libtrtx_hot.so > /dev/null 2>&1", shell=True)
seq = None
for f in glob.glob(tmp + "/*.cubin"):
    txt = subprocess.run(["nvdisasm", "-c", "-g", f], capture_output=True, text=True).stdout
    if mangled not in txt: continue
    cur, infn, s = None, False, []
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            infn = mangled in ln
            continue
        if ln.lstrip().startswith(".section") and infn and s: break
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/', ln): s.append(cur)
    if s: seq = s; break
extra = os.environ.get("NCU_EXTRA", "").split()  # e.g. NCU_EXTRA="--launch-skip 4 --launch-count 1" picks one launch
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kre}"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if '# Samples' in r: hdr = r; start = i + 1; break
j0 = hdr.index('# Samples'); je = hdr.index('Instructions Executed')
body = [r for r in rows[start:] if len(r) > j0]
n = len(seq)
samp, ex = [], []
for r in body[:n]:
    try: samp.append(float(r[j0] or 0)); ex.append(float(r[je] or 0))
    except ValueError: samp.append(0.0); ex.append(0.0)
print("sass instrs", n, "rows", len(body))
agg, aggx = collections.Counter(), collections.Counter()
for k, s_ in enumerate(samp):
    if seq[k]: agg[seq[k]] += s_; aggx[seq[k]] += ex[k]
tot, totx = sum(samp), sum(ex)
src = open(srcfile).read().splitlines()
base = os.path.basename(srcfile)
for (f, l), s_ in agg.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 25):
    txt = src[l - 1].strip()[:95] if f == base and l <= len(src) else f
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100 * s_ / tot, 100 * aggx[(f, l)] / totx, f, l, txt))
