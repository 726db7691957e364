"""Compact summary of an .ncu-rep (run where ncu is installed): key metrics + SASS hot regions."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp64_pred_on.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "?")[:90])
    for k in KEYS:
        if k in d:
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
            v = float(d[k] or 0)
            if v > 0.15:
                print(f"  stall {k.split('stalled_')[1].split('_per_issue')[0]:24s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
f = lambda r, k: float(r[ix[k]] or 0)
tot = sum(f(r, "Instructions Executed") for r in data)
thr = sum(f(r, "Thread Instructions Executed") for r in data)
print(f"SASS: {len(data)} instrs, warp-inst {tot:.3e}, avg threads {thr / tot:.2f}")
grp = int(sys.argv[2]) if len(sys.argv) > 2 else 80
for g in range(0, len(data), grp):
    blk = data[g:g + grp]
    n = sum(f(r, "Instructions Executed") for r in blk); t = sum(f(r, "Thread Instructions Executed") for r in blk)
    smp = sum(f(r, "# Samples") for r in blk)
    if n / tot > 0.008:
        c = collections.Counter((r[ix["Source"]].split() or [""])[0].split(".")[0].lstrip("@!P0123456789 ") for r in blk)
        print(f"  instr {g:4d}-{g + grp:4d}: {n / tot * 100:5.1f}% inst  avg thr {t / max(n, 1):5.1f}  samples {smp:8.0f}  {dict(c.most_common(5))}")
