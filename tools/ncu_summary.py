"""Summarise an `ncu -i X.ncu-rep --page raw --csv` dump: one line per captured launch with the counters the
roofline discussion needs (duration, tensor / XU / FMA / ALU / LSU pipe %, issue-slot %, DRAM bytes and % of peak,
L2 %, shared-memory bank conflicts, registers, smem, occupancy).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv [peak_GBs]
"""
import csv
import re
import sys

path = sys.argv[1]
peak_gbs = float(sys.argv[2]) if len(sys.argv) > 2 else 6531.9
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def get(r, name, default=float("nan")):
    i = ix.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return default
    u = units[i].split("/")[0]
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
             "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
    return v * scale


print(f"# {path}: ncu --set full --clock-control none, one line per captured launch; DRAM % is of the measured copy peak "
      f"({peak_gbs:.0f} GB/s)")
print(f"{'kernel':44s} {'grid':>10s} {'us':>7s} {'GHz':>5s} {'tensor%':>8s} {'xu%':>6s} {'fma%':>6s} {'alu%':>6s} "
      f"{'lsu%':>6s} {'issue%':>7s} {'dramRd MB':>10s} {'dramWr MB':>10s} {'GB/s':>7s} {'%peak':>6s} {'L2%':>5s} "
      f"{'smemConfl%':>10s} {'regs':>5s} {'smemKB':>7s} {'warps%':>7s}")
for r in data:
    name = r[ix["Kernel Name"]]
    short = re.sub(r"^void |\(anonymous namespace\)::|<unnamed>::|mfk::|\(.*$", "", name)[:44]
    grid = r[ix["Grid Size"]].replace(" ", "")
    us = get(r, "gpu__time_duration.sum")
    ghz = get(r, "sm__cycles_elapsed.max.per_second")
    rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / us / 1e3 if us == us and us > 0 else float("nan")
    wf = get(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    cf = get(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")
    confl = 100 * cf / wf if wf == wf and wf > 0 else 0.0
    print(f"{short:44s} {grid:>10s} {us:7.1f} {ghz:5.2f} "
          f"{get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):8.1f} "
          f"{get(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{get(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{get(r, 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{get(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{get(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{rd / 1e6:10.2f} {wr / 1e6:10.2f} {gbs:7.0f} {100 * gbs / peak_gbs:6.1f} "
          f"{get(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {confl:10.1f} "
          f"{get(r, 'launch__registers_per_thread'):5.0f} "
          f"{get(r, 'launch__shared_mem_per_block_dynamic') / 1e3 + get(r, 'launch__shared_mem_per_block_static', 0.0) / 1e3:7.1f} "
          f"{get(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):7.1f}")
