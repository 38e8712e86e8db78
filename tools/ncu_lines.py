"""Per-source-line view of an .ncu-rep: joins the SASS page of the report (address, stall samples, executed
instructions) with `nvdisasm -g` line information of the kernel's cubin.
Usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep ml_unigen_b200/csrc/decode_fast.o decode_fast_kernelILi32 [top]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout

# offset -> (file, line) for the requested kernel
line_of = {}
cur = None
infn = False
for ln in dis.splitlines():
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infn = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and infn:
        line_of[int(m.group(1), 16)] = cur

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
allrows = list(csv.reader(raw.splitlines()))
# the report may hold several launches: take the first section whose kernel name contains the plain function name
plain = re.sub(r"^_ZN\d+\w+?\d+|ILi\d+.*$", "", kern) if kern.startswith("_Z") else re.sub(r"ILi\d+.*$", "", kern)
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
sel = next((i for i in starts if plain in allrows[i][1]), starts[0])
end = next((j for j in starts if j > sel), len(allrows))
rows = allrows[sel:end]
hdr = rows[1]
ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iex]
base = int(data[0][ia], 16)
samples = collections.Counter()
execd = collections.Counter()
for r in data:
    off = int(r[ia], 16) - base
    key = line_of.get(off, ("?", 0))
    samples[key] += int(r[ismp] or 0)
    execd[key] += int(r[iex] or 0)
ts, te = sum(samples.values()), sum(execd.values())
print(f"total samples {ts}, executed warp instructions {te}")
print("| file:line | samples | % | executed | % |")
for key, s in samples.most_common(top):
    print(f"| {key[0]}:{key[1]} | {s} | {100 * s / ts:.1f} | {execd[key]} | {100 * execd[key] / te:.1f} |")
