"""profiles/traffic.json from ncu --set full captures (gpurun_out/<tag>_<cfg>.ncu-rep): DRAM bytes read + written by the captured
launch, the steps that launch ran, and the hash of the SASS of that kernel in the built library (bench.py reports
`roofline.traffic` only while the loaded kernel is, instruction for instruction, the captured one).  Run it with the library
the capture was taken from still built.  Configs without a capture under <tag> keep their entry.
usage: python tools/make_traffic.py r2 [cfg ...]"""
import csv, io, json, os, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import kernel_sass_sha

tag = sys.argv[1]
only = [c.lower() for c in sys.argv[2:]]
STEPS = {"c2": 20, "c2k": 20, "c3": 20, "c4": 1, "c4x": 1, "c5": 1}       # steps per captured launch (tools/profile_round.sh)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
path = "profiles/traffic.json"
out = json.load(open(path)) if os.path.exists(path) else {}
out.pop("src_sha", None)
out["what"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE captured launch (ncu --set full --clock-control none), divided "
               "by the steps that launch ran; sass_sha = hash of the captured kernel's SASS (bench.kernel_sass_sha)")
for cfg, k in STEPS.items():
    rep = f"gpurun_out/{tag}_{cfg}.ncu-rep"
    if not os.path.exists(rep) or (only and cfg not in only):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(name)
        tot += float(vals[i].replace(",", "")) * UNIT[units[i]]
    key = cfg.upper().replace("C2K", "C2k").replace("C4X", "C4x")
    out[key] = {
        "bytes_per_step": tot / k, "steps_in_captured_launch": k, "kernel": vals[hdr.index("Kernel Name")],
        "sass_sha": kernel_sass_sha(key), "capture": os.path.basename(rep),
        "duration_us_under_ncu": float(vals[hdr.index("gpu__time_duration.sum")].replace(",", "")) *
        {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[hdr.index("gpu__time_duration.sum")], 1.0)}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
