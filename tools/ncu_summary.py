"""Summarise a `ncu --set full` capture (one kernel launch) into a small JSON for profiles/:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.json [key=value ...]"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
extra = dict(a.split('=', 1) for a in sys.argv[3:])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: (u, v) for h, u, v in zip(hdr, units, vals)}


def num(name, scale_units=True):
    if name not in col:
        return None
    u, v = col[name]
    try:
        x = float(v.replace(',', ''))
    except ValueError:
        return None
    if scale_units:
        x *= {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12, 'ms': 1e3, 'msecond': 1e3, 'ns': 1e-3, 'nsecond': 1e-3,
              'second': 1e6, 's': 1e6}.get(u, 1.0)
    return x


summary = {
    'kernel_full': col.get('Kernel Name', ('', ''))[1],
    'grid': col.get('Grid Size', ('', ''))[1], 'block': col.get('Block Size', ('', ''))[1],
    'duration_us_under_ncu': num('gpu__time_duration.sum'),
    'dram_bytes_read': num('dram__bytes_read.sum'), 'dram_bytes_write': num('dram__bytes_write.sum'),
    'dram_throughput_pct': num('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
    'tensor_pipe_active_pct': num('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
    'tc_smem_read_pipe_pct': num('l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
    'sm_throughput_pct': num('sm__throughput.avg.pct_of_peak_sustained_elapsed'),
    'l1tex_throughput_pct': num('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'),
    'lts_throughput_pct': num('lts__throughput.avg.pct_of_peak_sustained_elapsed'),
    'warps_active_pct': num('sm__warps_active.avg.pct_of_peak_sustained_active'),
    'registers_per_thread': num('launch__registers_per_thread'),
    'ctas_per_sm_limit_smem': num('launch__occupancy_limit_shared_mem'),
    'ctas_per_sm_limit_regs': num('launch__occupancy_limit_registers'),
}
summary.update(extra)
json.dump(summary, open(out, 'w'), indent=1)
print(json.dumps(summary, indent=1))
