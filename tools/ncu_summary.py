#!/usr/bin/env python
"""Print the handful of counters DESIGN.md / profiles/README.md quote from an .ncu-rep (needs `ncu` on PATH)."""
import csv
import re
import subprocess
import sys

PAT = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__t_sector_hit_rate\.pct|"
    r"l1tex__m_xbar2l1tex_read_bytes\.sum(\.per_second)?|lts__t_bytes\.sum(\.per_second)?|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|"
    r"launch__occupancy_limit_(registers|warps|shared_mem|blocks)|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"smsp__inst_executed\.sum|smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|lg_throttle|mio_throttle|"
    r"math_pipe_throttle|wait|not_selected|barrier|tex_throttle)_per_issue_active\.ratio)$")


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            print(f"== {path} :: {name}")
            for h, u, v in zip(hdr, units, vals):
                if PAT.match(h):
                    print(f"  {h:90s} {v} {u}")


if __name__ == "__main__":
    main()
