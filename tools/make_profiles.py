"""Regenerate the round's judged profile summaries under profiles/ from the scratch files in gpurun_out/.

    python tools/make_profiles.py <bench.json> <launches.csv> <walk.ncu-rep> [round tag, default r2]
"""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bench, launches, rep = sys.argv[1:4]
TAG = sys.argv[4] if len(sys.argv) > 4 else "r2"
line = json.loads(open(bench).read().strip().splitlines()[-1])
with open(os.path.join(ROOT, "profiles", f"{TAG}_bench_final.json"), "w") as f:
    f.write(json.dumps(line) + "\n")
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
each = collections.defaultdict(list)
for r in rows:
    name = r[4].split("(")[0].replace("wgrt::<unnamed>::", "").replace("void ", "")
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[-1])
    each[name].append(float(r[-1]))
tot = sum(a[1] for a in agg.values())
out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-reference-gpu --no-partitioned --no-legacy",
       "(per-launch times under ncu are serialised and cold-cache; what must agree with bench.py is the SHARE of the step.",
       " One step of bench.py = 17 kernel launches (+ two 32-byte device-to-device copies into constant memory): region_hash, bbox, rowmask, coarse, fine, atlas, atlas2, zone_reset, zone_level1, zone_collect2,",
       " zone_number, zone_write1, zone_write2, zone_trans (all no-ops after the first launch of a geometry), pick_tile_warp, walk_warp<0,0,1>,",
       " walk_redo (the near-tie rays; usually none).  walk_strict<1> / walk_warp<1,0,1> / fma_peak are the counter replay and the roofline",
       " probes, outside the timed region.)",
       f"{'kernel':45s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg ms':>9s}"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k[:45]:45s} {n:8d} {t / 1e6:10.3f} {t / tot * 100:6.2f}% {t / n / 1e6:9.4f}")
import statistics
step = sum(statistics.median(each[k]) for k in each if k.startswith(("region_", "zone_", "pick_tile", "walk_redo")))
walk = statistics.median(next(v for k, v in each.items() if k.startswith("walk_warp_kernel<0, 0")))
first = sum(each[k][0] for k in each if k.startswith(("region_", "zone_")))
out.append(f"per step (median launch of each kernel): walk_warp {walk / 1e6:.3f} ms of {(walk + step) / 1e6:.3f} ms = "
           f"{walk / (walk + step) * 100:.2f} % (bench.py: {line['ms_per_step']:.2f} ms per step); "
           f"the region index + atlas build of the FIRST launch of a geometry: {first / 1e6:.2f} ms")
open(os.path.join(ROOT, "profiles", f"{TAG}_launch_shares_final.txt"), "w").write("\n".join(out) + "\n")
import shutil
shutil.copy(launches, os.path.join(ROOT, "profiles", f"{TAG}_launches_final_bench_steps2.csv"))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw))); d = dict(zip(rr[0], rr[2]))
g = lambda k: float(d[k].replace(",", ""))
units = dict(zip(rr[0], rr[1]))
def byts(k):
    v = g(k); u = units[k]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
rd, wr, ms = byts("dram__bytes_read.sum"), byts("dram__bytes_write.sum"), g("gpu__time_duration.sum") * (1e-3 if units["gpu__time_duration.sum"] == "us" else 1)
traffic = {"kernel": "walk_warp_kernel", "source": f"ncu --set full --clock-control none, {os.path.basename(rep)} (timed C2 launch of bench.py, 112.5M rays, explicit ray arrays)",
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "gpu_time_duration_ms": ms,
           "threads_per_warp_inst": g("smsp__thread_inst_executed_per_inst_executed.ratio"),
           "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "fp64_pipe_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           "registers": g("launch__registers_per_thread"), "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active")}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "walk_warp_traffic.json"), "w"), indent=1)
tools = os.path.join(ROOT, "tools")
s1 = subprocess.run([sys.executable, os.path.join(tools, "ncu_summary.py"), rep, "100"], capture_output=True, text=True).stdout
s2 = subprocess.run([sys.executable, os.path.join(tools, "ncu_lines.py"), rep, "40"], capture_output=True, text=True).stdout
hdr = ("ncu --set full --clock-control none --import-source on -k regex:walk_warp -s 3 -c 1: python bench.py --steps 2 --warmup 3 "
       "(timed launch of C2, 112.5 M rays, explicit ray arrays)\n")
der = (f"\nDerived: DRAM {(rd + wr) / 1e9:.2f} GB / {ms:.2f} ms = {(rd + wr) / ms / 1e6:.0f} GB/s = {(rd + wr) / ms / 1e6 / 6549.1 * 100:.1f} % of the measured 6549 GB/s copy peak;\n"
       f"FP64 pipe {traffic['fp64_pipe_pct']:.0f} % of peak, issue slots {traffic['issue_active_pct']:.0f} % busy, {traffic['threads_per_warp_inst']:.1f} of 32 lanes per warp "
       f"instruction, {traffic['warps_active_pct'] * 0.64:.0f} of 64 warp slots.\n\n")
open(os.path.join(ROOT, "profiles", f"{TAG}_walk_warp_ncu_summary.txt"), "w").write(hdr + s1 + der + s2)
print(out[-1]); print(der)
