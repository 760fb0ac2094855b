"""Per-barrier wait statistics of a warp-specialised kernel from an ncu report (source page):
for every mbarrier try_wait, how often it was executed and how often it had to retry -- shows which
role (producer / MMA issuer / epilogue) waits on which.   python tools/ncu_waits.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[start], rows[start + 1:]
i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
prev = None
for n, r in enumerate(data):
    if len(r) <= i_ex:
        continue
    src = r[i_src]
    if "SYNCS.PHASECHK" in src or "UTCHMMA" in src or "UTMALDG" in src or "LDTM" in src:
        if int(r[i_ex] or 0) == 0:
            continue
        print(f"{n:5d} samples={r[i_s]:>6s} executed={r[i_ex]:>10s}  {src.strip()[:90]}")
