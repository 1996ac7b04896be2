"""Summarise an ncu --set full capture of the transport kernel into a small JSON for profiles/.
    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/name.json "note"
"""
import csv
import json
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_sectors_srcunit_tex_op_atom.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_red.sum', 'l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'launch__shared_mem_per_block_dynamic']


def main(rep, out, note):
    # a .csv argument is the raw page already exported on the GPU box (scripts/gpu_ncu_cmd.sh keeps only the CSV pages)
    raw = open(rep).read() if rep.endswith('.csv') else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    h, u, v = r[0], r[1], r[2]
    d = {}
    for i, k in enumerate(h):
        if k in WANT or ('pcsamp_warps_issue_stalled' in k and 'not_issued' not in k) or k == 'Kernel Name':
            d[k] = {'unit': u[i], 'value': v[i]}
    d['_note'] = note
    json.dump(d, open(out, 'w'), indent=1)
    print(out, len(d))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else '')
