"""Reads the csv exports of an ncu capture (profiles/ or gpurun_out/): key raw metrics + per-instruction hot spots.
   python profiles/hot.py gpurun_out/r02_c3_fused [min_sample_pct]"""
import csv, sys
base = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
rows = list(csv.reader(open(base + "_raw.csv")))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:110])
    for k in WANT:
        if k in d:
            print(f"   {k:85s}{d[k]:>16s} {units[hdr.index(k)]}")
# source page: one block per kernel, starting with a "Kernel Name" row
rows = list(csv.reader(open(base + "_source.csv")))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
for b in blocks:
    ix = {h: i for i, h in enumerate(b["hdr"])}
    tot = sum(int(r[ix["# Samples"]] or 0) for r in b["rows"]) or 1
    print("\n== source:", b["name"][:100], "samples", tot)
    print(f"{'sass':72s}{'samp%':>6s}{'inst':>9s}{'tagreq':>9s}{'shwave':>9s}{'l2sect':>10s} lsb ssb mio bar math")
    for r in b["rows"]:
        s = int(r[ix["# Samples"]] or 0)
        tag = int(r[ix["L1 Tag Requests Global"]] or 0)
        sh = int(r[ix["L1 Wavefronts Shared"]] or 0)
        if 100 * s / tot >= thr or tag > 0 or (sh > 0 and len(sys.argv) > 3):
            print(f"{r[ix['Source']].strip()[:72]:72s}{100*s/tot:6.2f}{int(r[ix['Instructions Executed']]):9d}{tag:9d}{sh:9d}"
                  f"{int(r[ix['L2 Theoretical Sectors Global']] or 0):10d} {r[ix['stall_long_sb']]:>4s}{r[ix['stall_short_sb']]:>5s}"
                  f"{r[ix['stall_mio']]:>5s}{r[ix['stall_barrier']]:>5s}{r[ix['stall_math']]:>5s}")
    gt = sum(int(r[ix["L1 Tag Requests Global"]] or 0) for r in b["rows"])
    st = sum(int(r[ix["L1 Wavefronts Shared"]] or 0) for r in b["rows"])
    print(f"   total global tag requests {gt}   shared wavefronts {st}")
