"""Key metrics + top stall lines of a .ncu-rep (ncu --set full --import-source on)."""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread ', 'launch__grid_size', 'launch__block_size', 'lts__throughput.avg.pct', 'sm__throughput.avg.pct',
        'l1tex__throughput.avg.pct', 'launch__occupancy_limit', 'sm__cycles_elapsed.avg ', 'smsp__cycles_active.avg ',
        'smsp__average_warps_issue_stalled', 'dram__bytes_read.sum.per_second', 'launch__waves_per_multiprocessor']

def main(path, top=14):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:3]:
        name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''
        print('kernel:', name[:120])
        stalls = []
        for h, u, v in zip(hdr, units, vals):
            hh = h + ' '
            if 'issue_stalled' in h and 'per_issue_active' in h:
                try: stalls.append((float(v), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError: pass
            elif any(k in hh for k in KEYS) and 'issue_stalled' not in h:
                print(f'  {h} [{u}] = {v}')
        stalls.sort(reverse=True)
        print('  stall reasons (warps per issue):', ', '.join(f'{n}={v:.2f}' for v, n in stalls[:7]))
    src = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
    if not hi:
        return
    h = rows[hi[0]]
    si = h.index('# Samples')
    data = [r for r in rows[hi[0] + 1:] if len(r) > si]
    def num(x):
        try: return float(x)
        except ValueError: return 0.0
    tot = sum(num(r[si]) for r in data) or 1
    data.sort(key=lambda r: -num(r[si]))
    print(f'  SASS lines: {len(data)}, samples {tot:.0f}; top:')
    for r in data[:top]:
        print(f'   {100 * num(r[si]) / tot:5.1f}%  {r[1][:100]}')

if __name__ == '__main__':
    for p in sys.argv[1:]:
        main(p)
        print()
