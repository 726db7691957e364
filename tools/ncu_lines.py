"""Per-CUDA-source-line view of an .ncu-rep: warp instructions, avg threads, stall samples."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
out, fname, hdr = [], "?", None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 10 or not r[0]:
        continue
    def g(k):
        try:
            return float(r[hdr[k]] or 0)
        except ValueError:
            return 0.0
    out.append((g("Instructions Executed"), g("Thread Instructions Executed"), g("# Samples"), fname, r[0], r[1].strip()))
tot = sum(o[0] for o in out); thr = sum(o[1] for o in out); smp = sum(o[2] for o in out)
print(f"total warp-inst {tot:.3e}  avg threads {thr / tot:.2f}  samples {smp:.0f}")
out.sort(key=lambda o: -o[2])
print("  %inst  %smp  thr  file:line  source")
for n, t, s, f, ln, text in out[:top]:
    print(f"  {n / tot * 100:5.2f} {s / smp * 100:5.2f} {t / max(n, 1):5.1f}  {f}:{ln}  {text[:110]}")
