"""Summarise an ncu report: key raw metrics + the hottest SASS lines (stall sampling) of each kernel instance."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-70s %s %s" % (w, r[i], units[i]))
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; idx = {k: i for i, k in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h) and r[idx['# Samples']].isdigit()]
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print("stall sampling, first kernel instance: %d samples; hottest SASS lines:" % tot)
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    n = int(r[idx['# Samples']] or 0)
    st = sorted([(k[6:], int(r[idx[k]] or 0)) for k in stalls if int(r[idx[k]] or 0) > 0], key=lambda kv: -kv[1])[:2]
    print("%6d %5.1f%%  exec %9s  %-64s %s" % (n, 100.0 * n / max(tot, 1), r[idx['Instructions Executed']], r[idx['Source']][:64], st))
