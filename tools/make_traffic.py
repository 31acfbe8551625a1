"""profiles/traffic.json from the round's ncu --set full captures (gpurun_out/<tag>_<cfg>.ncu-rep): DRAM bytes read + written by
the captured launch, the steps that launch ran, and the hash of the CUDA sources the capture belongs to (bench.py reports
`roofline.traffic` only while the sources still hash to it).   usage: python tools/make_traffic.py r2"""
import csv, io, json, os, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_source_sha

tag = sys.argv[1]
STEPS = {"c2": 20, "c2k": 20, "c3": 20, "c4": 1, "c4x": 1, "c5": 1}       # steps per captured launch (tools/profile_round.sh)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {"src_sha": kernel_source_sha(), "what": "dram__bytes_read.sum + dram__bytes_write.sum of ONE captured launch "
       "(ncu --set full --clock-control none), divided by the steps that launch ran"}
for cfg, k in STEPS.items():
    rep = f"gpurun_out/{tag}_{cfg}.ncu-rep"
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        tot += float(vals[i].replace(",", "")) * UNIT[units[i]]
    out[cfg.upper().replace("C2K", "C2k").replace("C4X", "C4x")] = {
        "bytes_per_step": tot / k, "steps_in_captured_launch": k, "kernel": vals[hdr.index("Kernel Name")],
        "duration_us_under_ncu": float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
        {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)}
json.dump(out, open("profiles/traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
