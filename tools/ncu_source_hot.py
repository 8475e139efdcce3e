"""Top source lines of one kernel of an .ncu-rep by warp-stall samples (ncu --page source, -lineinfo builds).
usage: python tools/ncu_source_hot.py REP.ncu-rep KERNEL_REGEX [N]"""
import subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx,
                      "--print-source=cuda,sass"], capture_output=True, text=True).stdout
rows, cur, hdr = [], None, None
for line in raw.splitlines():
    if not line.startswith('"'):
        continue
    r = line[1:-1].split('","')
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].isdigit() and len(r) >= len(hdr):
        extra = len(r) - len(hdr)           # quotes / commas inside the source text
        src = '","'.join(r[1:2 + extra])
        r = [r[0], src] + r[2 + extra:]
        i_s, i_i = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        try:
            rows.append((int(r[i_s] or 0), cur, int(r[0]), src.strip()[:120], int(r[i_i] or 0)))
        except ValueError:
            pass
tot = sum(x[0] for x in rows) or 1
print("total samples", tot)
for s, f, ln, src, ins in sorted(rows, reverse=True)[:n]:
    print(f"{100.0 * s / tot:5.1f}%  {f}:{ln:<5d} inst={ins:<10d} {src}")
