"""Summarises an `ncu --set full --page raw --csv` export of the conv kernels of one inference step (dev / evidence tool):
per launch the duration, tcgen05 (utchmma) pipe %, shared-memory / L1 / L2 / DRAM figures, and the time-weighted means.
python tools/ncu_conv_summary.py gpurun_out/r2_conv_ncu_raw.csv > profiles/r2_conv_ncu_summary.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
COLS = {"us": "gpu__time_duration.sum",
        "utchmma%": "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "tc_smem_wf%": "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lsu_wf%": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex%": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts%": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram_rd_MB": "dram__bytes_read.sum", "dram_wr_MB": "dram__bytes_write.sum",
        "grid": "launch__grid_size", "block": "launch__block_size", "regs": "launch__registers_per_thread"}
idx = {k: (hdr.index(v) if v in hdr else None) for k, v in COLS.items()}


def val(r, k):
    i = idx[k]
    try:
        return float(r[i].replace(",", "")) if i is not None else float("nan")
    except ValueError:
        return float("nan")


data = rows[2:]
t = [val(r, "us") for r in data]
mbs = [val(r, "dram_rd_MB") + val(r, "dram_wr_MB") for r in data]
# the step starts at the stem: it and the next conv (P2) are the two largest-traffic launches and are adjacent
start = int(sys.argv[2]) if len(sys.argv) > 2 else max(range(len(t)), key=lambda i: mbs[i] + mbs[(i + 1) % len(t)])
order = list(range(start, len(data))) + list(range(start))
print(f"source: {sys.argv[1]} ({len(data)} conv_tc_kernel launches of one step, ncu --set full --clock-control none; cold-cache, serialised)")
print("  #      us  grid blk regs utchmma% tc-smem-wf%  lsu-wf%  l1tex%   lts%  dram MB (rd+wr)  DRAM GB/s")
tot = 0.0
acc = {k: 0.0 for k in ("utchmma%", "tc_smem_wf%", "l1tex%", "lts%")}
bytes_tot = 0.0
for n, i in enumerate(order):
    r = data[i]
    us = val(r, "us")
    mb = val(r, "dram_rd_MB") + val(r, "dram_wr_MB")
    tot += us
    bytes_tot += mb
    for k in acc:
        acc[k] += us * val(r, k)
    print(f"{n:3d} {us:7.1f} {int(val(r, 'grid')):5d} {int(val(r, 'block')):3d} {int(val(r, 'regs')):4d} {val(r, 'utchmma%'):8.1f} {val(r, 'tc_smem_wf%'):10.1f} "
          f"{val(r, 'lsu_wf%'):8.1f} {val(r, 'l1tex%'):7.1f} {val(r, 'lts%'):6.1f} {mb:12.1f} {mb / us * 1e3:10.0f}")
print(f"total {tot:.1f} us; time-weighted: " + ", ".join(f"{k} {v / tot:.1f}" for k, v in acc.items()) +
      f"; DRAM traffic {bytes_tot:.0f} MB per step = {bytes_tot / len(data):.1f} MB per launch, {bytes_tot / tot * 1e3:.0f} GB/s")
