"""Top source lines of an ncu report by warp-stall samples (with the dominant stall reasons)."""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = hdr = None
per = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter()])
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if r[0].isdigit() and hdr:
        k = (cur, int(r[0]))
        try:
            per[k][0] += float(r[6]); per[k][1] += float(r[7])
        except Exception:
            pass
        for i, name in enumerate(hdr):
            if name.startswith("stall_") and "Not Issued" not in name:
                try:
                    per[k][2][name] += float(r[i])
                except Exception:
                    pass
tot = sum(v[0] for v in per.values())
for k, v in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    why = ", ".join(f"{n[6:]}:{int(c)}" for n, c in v[2].most_common(3))
    print(f"{k[0]:20s} {k[1]:5d}  samples {v[0]:6.0f} ({100 * v[0] / tot:4.1f}%)  instr {v[1]:9.0f}  {why}")
