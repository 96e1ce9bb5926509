"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md argues with."""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'smsp__sass_thread_inst_executed_op_fp32_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_per_inst_issued.ratio', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_lg.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum']


def main(path):
    raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        print(f"## {d['Kernel Name'][:70]} (id {d.get('ID', '?')})\n\n| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in d:
                print(f"| {k} | {d[k]} | {u[k]} |")
        st = {k: float(v) for k, v in d.items() if 'issue_stalled' in k and k.endswith('_per_warp_active.pct') and v}
        for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]:
            print(f"| {k} | {v:.2f} | % |")
        print()


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
