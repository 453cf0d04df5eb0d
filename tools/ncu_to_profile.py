"""Turn `ncu -i X.ncu-rep --page raw --csv` into the per-kernel JSON summary committed under profiles/ and update
profiles/traffic.json (DRAM bytes per launch of each kernel, used by bench.py's roofline.traffic).

usage: python tools/ncu_to_profile.py raw.csv profiles/rNN_name.json [sedov<side> to record traffic]"""
import csv
import json
import re
import sys
from pathlib import Path

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
PHASE = {"CapsStd": "block_search", "CapsBig": "block_search_overflow", "XMassOp": "xmass", "GradhOp": "ve_def_gradh", "IadOp": "iad_divv_curlv",
         "AvOp": "av_switches", "MomentumOp": "momentum_energy", "eosKernel": "eos"}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def main(raw, out, traffic_tag=None):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    res, traffic = [], {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        d = {"kernel": re.sub(r"^void ", "", name)[:60]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = f"{r[i]} {units[i]}".strip()
        res.append(d)
        for key, ph in PHASE.items():
            if key in name and "dram__bytes_read.sum" in hdr:
                ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                traffic[ph] = float(r[ir]) * UNIT.get(units[ir], 1.0) + float(r[iw]) * UNIT.get(units[iw], 1.0)
                if "smsp__inst_executed.sum" in hdr:  # warp instructions per launch (bench.py: issue-slot bound)
                    traffic["inst:" + ph] = float(r[hdr.index("smsp__inst_executed.sum")])
    Path(out).write_text(json.dumps(res, indent=1))
    if traffic_tag:
        tf = Path(out).parent / "traffic.json"
        t = json.loads(tf.read_text()) if tf.exists() else {}
        for ph, b in traffic.items():
            t[f"{ph}@{traffic_tag}"] = b
        tf.write_text(json.dumps(t, indent=1, sort_keys=True))
    print(json.dumps(traffic))


if __name__ == "__main__":
    main(*sys.argv[1:4])
