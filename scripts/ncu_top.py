"""Summarise an .ncu-rep: headline metrics + top stall locations (run where ncu is installed)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum",
        "sm__pipe_fp32_cycles_active", "smsp__cycles_active.avg"]
for h, u, v in zip(hdr, units, vals):
    if any(h.endswith(w) or h == w for w in want):
        print(f"{h} [{u}] = {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
print('kernel:', rows[0][1][:90], ' total samples', tot)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
print('stall totals:', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:ntop]:
    st = {h: int(r[ix[h]] or 0) for h in stalls}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(r[ix['Address']][-5:], r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[ix['Source']][:64].ljust(64), main)
