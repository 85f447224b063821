"""Prints the key metrics + wait-loop attribution of an ncu report of the fused MLP kernel (reads .ncu-rep here, no GPU)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
for k in ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
          "sm__cycles_elapsed.max.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
          "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "launch__registers_per_thread", "launch__cluster_size",
          "launch__grid_size", "launch__block_size"]:
    if k in d:
        print(f"{k:88s} {d[k][0]:10s} {d[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
names = {"0x20": "W_EMPTY (producer)", "0x28": "W_EMPTY", "0x30": "W_EMPTY", "0x38": "W_EMPTY", "0x50": "ACC_FULL (epilogue)",
         "0x58": "ACC_FULL (epilogue)", "38040": "A_READY[0] (MMA)", "38048": "A_READY[1] (MMA)"}
for i, r in enumerate(data):
    s_ = r[ix["Source"]]
    if "TRYWAIT" in s_:
        acc = 0; j = i
        while j < len(data) and j < i + 12:
            acc += int(data[j][ix["# Samples"]] or 0)
            if "BRA" in data[j][ix["Source"]] and j > i: break
            j += 1
        if acc / tot > 0.002:
            print(f"  wait loop {100 * acc / tot:6.2f}%  {s_.strip()[:80]}")
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]
for r in top:
    s_ = int(r[ix["# Samples"]])
    print(f"{s_:7d} {100 * s_ / tot:5.1f}% lsb={r[ix['stall_long_sb']]:>6s} wait={r[ix['stall_wait']]:>5s} ssb={r[ix['stall_short_sb']]:>5s} mio={r[ix['stall_mio']]:>5s} {r[ix['Source']][:80]}")
