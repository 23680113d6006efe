"""Aggregate an ncu report's warp-stall samples / executed instructions by CUDA source line.
usage: python profiles/src_hotspots.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0] != "" and len(r) > 8:
        try:
            agg.append((cur, r[0], r[1], float(r[6] or 0), float(r[7] or 0), float(r[8] or 0)))
        except ValueError:
            pass
ts = sum(a[3] for a in agg) or 1; ti = sum(a[4] for a in agg) or 1; tt = sum(a[5] for a in agg) or 1
print(f"total samples {ts:.0f} warp-inst {ti:.0f} thread-inst {tt:.0f} avg active threads {tt/ti:.1f}")
for a in sorted(agg, key=lambda x: -x[3])[:top_n]:
    print(f"{a[0]:20s} {a[1]:>5s} samp {a[3]/ts*100:5.1f}% inst {a[4]/ti*100:5.1f}% thr/inst {a[5]/max(a[4],1):5.1f} {a[2].strip()[:100]}")
