"""Condense an `ncu --page raw --csv` export into the per-kernel summary committed under
profiles/ (dev tool): python tools/ncu_summary.py <raw.csv> <out_summary.csv> <out_traffic.json>"""
import csv, json, sys
KEEP = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__throughput.avg.pct_of_peak_sustained_active',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
    'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
name_col = hdr.index('Kernel Name')
kern = [r[name_col][:48] for r in data]
with open(sys.argv[2], 'w', newline='') as fp:
    w = csv.writer(fp)
    w.writerow(['metric'] + kern)
    for m in KEEP:
        if m in hdr:
            c = hdr.index(m)
            w.writerow([m] + [f"{r[c]} {units[c]}".strip() for r in data])


def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]


if len(sys.argv) > 3:
    cr, cw = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
    out = {"source": f"ncu --set full --clock-control none ({sys.argv[1]}), per launch: "
                     "dram__bytes_read.sum + dram__bytes_write.sum"}
    for r in data:
        nm = r[name_col]
        base = nm.split('(')[0].split('<')[0].replace('void ', '').strip()
        key = {'k_eamz_force': 'k_eamz_force<f64>', 'k_eamz_rho': 'k_eamz_rho<f64>'}.get(base, base)
        if key in out:          # several launches of one kernel: keep the first
            continue
        out[key] = int(to_bytes(r[cr], units[cr]) + to_bytes(r[cw], units[cw]))
    json.dump(out, open(sys.argv[3], 'w'), indent=1)
