"""Summarise an `ncu --set full` report (run here, no GPU needed): per kernel, the metrics the roofline and the
DESIGN.md tables quote, plus the top stall sites of the source page.

    python tools/ncu_metrics.py gpurun_out/prof.ncu-rep > profiles/r01_prof_x.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread ",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_tmem_ldt.sum ",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active", "smsp__average_warps_issue_stalled_wait_per_issue_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active", "smsp__average_warps_issue_stalled_barrier_per_issue_active",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(path):
    raw = list(csv.reader(io.StringIO(run(["-i", path, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    print(f"# {path}")
    tables = None
    for n, row in enumerate(raw[2:], 1):
        d = dict(zip(hdr, row))
        print(f"\n## launch {n}: {d.get('Kernel Name', '?')[:100]}")
        for h, u in zip(hdr, units):
            if any(h.startswith(k.strip()) and (k.endswith(' ') and h == k.strip() or not k.endswith(' ')) for k in KEYS):
                print(f"{h:90s} {d[h]:>16s} {u}")
        if tables is None:           # the source page of the whole report, split into one table per launch ("Kernel Name" rows)
            tables = []
            for r in csv.reader(io.StringIO(run(["-i", path, "--page", "source", "--csv"]))):
                if r and r[0] == "Kernel Name":
                    tables.append([r[1] if len(r) > 1 else "", []])
                elif tables:
                    tables[-1][1].append(r)
        # the source page holds `per` tables per launch in launch order (with --import-source: the SASS view, then the CUDA-C view)
        per = max(1, len(tables) // max(1, len(raw) - 2))
        pick = tables[(n - 1) * per] if (n - 1) * per < len(tables) else None
        if pick is None or len(pick[1]) < 2:
            continue
        sh = pick[1][0]
        ix = {h: i for i, h in enumerate(sh)}
        data = [r for r in pick[1][1:] if len(r) == len(sh)]
        if '# Samples' not in ix or not data:
            continue
        tot = sum(int(r[ix['# Samples']]) for r in data) or 1
        print(f"-- top stall sites ({tot} samples, {sum(int(r[ix['Instructions Executed']]) for r in data)} warp instructions)")
        for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:12]:
            st = {k.replace('stall_', ''): int(r[ix[k]]) for k in sh if k.startswith('stall_') and 'Not Issued' not in k and int(r[ix[k]]) > 0}
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            print(f"   {100.0 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:60]:60s} {top}")


if __name__ == "__main__":
    main(sys.argv[1])
