"""Turns .ncu-rep captures into the small text artefacts committed under profiles/ (run in the build container: ncu reads
reports without a GPU).   python scripts/ncu_export.py <report.ncu-rep> <out_prefix> [kernel-regex]
Writes <out_prefix>_raw.csv (ncu --page raw --csv, one row per captured launch, every metric of the capture) and prints the
metrics the profile notes quote.  With --traffic <samples per launch> it also writes profiles/r02_mlp_traffic.json, which
bench.py reads for roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum per sample of the fused MLP kernel)."""
import csv, io, json, os, subprocess, sys

argv = list(sys.argv[1:])
if "--traffic" in argv:
    i = argv.index("--traffic"); del argv[i:i + 2]
args = [a for a in argv if not a.startswith("--")]
rep, prefix = args[0], args[1]
regex = args[2] if len(args) > 2 else None
cmd = ["ncu", "-i", rep, "--page", "raw", "--csv"]
if regex:
    cmd += ["--kernel-name", f"regex:{regex}"]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
open(prefix + "_raw.csv", "w").write(raw)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size"]
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)
for r in rows[2:]:
    d = {h: (u, v) for h, u, v in zip(hdr, units, r)}
    print(d["Kernel Name"][1][:80])
    for k in KEYS:
        if k in d:
            print(f"    {k:72s} {d[k][1]:>16s} {d[k][0]}")
if "--traffic" in sys.argv:
    samples = float(sys.argv[sys.argv.index("--traffic") + 1])
    d = {h: (u, v) for h, u, v in zip(hdr, units, rows[2])}
    tot = to_bytes(d["dram__bytes_read.sum"][1], d["dram__bytes_read.sum"][0]) + to_bytes(d["dram__bytes_write.sum"][1], d["dram__bytes_write.sum"][0])
    out = {"dram_bytes_per_sample": tot / samples, "dram_bytes_read": to_bytes(d["dram__bytes_read.sum"][1], d["dram__bytes_read.sum"][0]),
           "dram_bytes_write": to_bytes(d["dram__bytes_write.sum"][1], d["dram__bytes_write.sum"][0]), "samples_per_launch": samples,
           "kernel": d["Kernel Name"][1], "source": os.path.basename(prefix) + "_raw.csv (ncu --set full --clock-control none, scripts/gpu_r2_profile.sh)"}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(prefix)), "r02_mlp_traffic.json"), "w"), indent=1)
    print("traffic:", out)
