"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.txt [note]
"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_read\.sum|dram__bytes_write\.sum|"
    r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__cycles_elapsed\.avg|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|"
    r"launch__grid_size|launch__block_size|launch__shared_mem_per_block_dynamic|"
    r"launch__occupancy_limit_\w+|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum|"
    r"l1tex__t_bytes\.sum|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_sectors_srcunit_tex\.sum|sm__inst_executed_pipe_tma\.sum|smsp__inst_executed\.sum|sm__inst_executed_pipe_uniform\.sum|"
    r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|"
    r"smsp__cycles_active\.avg)$")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s\n" % rep)
        if note:
            f.write("# %s\n" % note)
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("\n== launch %s: %s  grid %s block %s\n" % (d.get("ID"), d.get("Kernel Name", "")[:110],
                                                               d.get("Grid Size"), d.get("Block Size")))
            stalls = []
            for name, unit, val in zip(hdr, units, r):
                if not KEEP.match(name):
                    continue
                if name.startswith("smsp__average_warps_issue_stalled"):
                    try:
                        stalls.append((float(val), name))
                    except ValueError:
                        pass
                    continue
                f.write("%-78s %12s %s\n" % (name, val, unit))
            for v, name in sorted(stalls, reverse=True)[:6]:
                f.write("%-78s %12.3f\n" % (name, v))


if __name__ == "__main__":
    main()
