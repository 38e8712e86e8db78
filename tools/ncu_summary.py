"""Summarise an .ncu-rep (ncu --set full) into the markdown tables committed under profiles/.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/xyz.md"""
import csv, subprocess, sys, collections

rep, title = sys.argv[1], sys.argv[2]
kfilter = sys.argv[3] if len(sys.argv) > 3 else None          # substring of the kernel name (reports with several launches)
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
kcol = hdr.index("Kernel Name")
vals = next((r for r in rows[2:] if kfilter is None or kfilter in r[kcol]), rows[2])
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# {title}\n")
print(f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`, one launch; numbers under the profiler are not bench values).\n")
print("| metric | unit | value |\n|---|---|---|")
for h, u, v in zip(hdr, units, vals):
    if h in want:
        print(f"| {h} | {u} | {v[:120]} |")
print("\n## Warp stall reasons (average warps stalled per issue-active cycle)\n\n| reason | value |\n|---|---|")
st = [(h, float(v)) for h, v in zip(hdr, vals) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for h, v in sorted(st, key=lambda kv: -kv[1])[:10]:
    print(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"]
sel = next((i for i in starts if kfilter is None or kfilter in srows[i][1]), starts[0])
srows = srows[sel:next((j for j in starts if j > sel), len(srows))]
h2 = srows[1]
isrc, iex, ismp = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
data = [r for r in srows[2:] if len(r) > iex]
tot = sum(int(r[iex] or 0) for r in data)
tots = sum(int(r[ismp] or 0) for r in data)
c = collections.Counter(); cs = collections.Counter()
for r in data:
    s = r[isrc]
    op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
    c[op] += int(r[iex] or 0); cs[op] += int(r[ismp] or 0)
print(f"\n## Executed warp instructions by opcode (total {tot}; sampled stalls {tots})\n\n| opcode | executed | share | share of stall samples |\n|---|---:|---:|---:|")
for op, e in c.most_common(18):
    print(f"| {op} | {e} | {100 * e / tot:.1f}% | {100 * cs[op] / max(tots, 1):.1f}% |")
print("\n## Hottest SASS lines by stall samples\n\n| SASS | samples | executed |\n|---|---:|---:|")
for r in sorted(data, key=lambda r: -int(r[ismp] or 0))[:12]:
    print(f"| `{r[isrc][:80]}` | {r[ismp]} | {r[iex]} |")
