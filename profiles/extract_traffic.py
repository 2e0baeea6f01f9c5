#!/usr/bin/env python
"""profiles/extract_traffic.py <ncu-rep> <out.json>
DRAM traffic (ncu dram__bytes_read.sum + dram__bytes_write.sum) of the captured greedy_loop_kernel / scan_kernel launches
of tools/ncu_traffic_case.py next to their algorithmic bytes (4 N S per pass; with the float16 pre-filter the kernel
moves about half of that).  bench.py reads the ratio from it
(roofline.traffic)."""
import csv
import json
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
# (kernel substring, S, float16 pre-filter, algorithmic bytes per launch) in launch order of tools/ncu_traffic_case.py
A7, A6 = 4.*10_000_000*512, 4.*1_000_000*256
EXPECT = [('greedy_loop_kernel', 512, True, A7*5), ('omp_loop_kernel', 512, True, A7*2), ('greedy_loop_kernel', 512, False, A7*5),
          ('omp_loop_kernel', 512, False, A7*2), ('scan_kernel', 512, False, A7), ('scan_kernel', 512, False, A7),
          ('greedy_loop_kernel', 256, True, A6*5), ('greedy_loop_kernel', 256, False, A6*5)]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if 'exact_scan_kernel' not in r[rows[0].index('Kernel Name')]]
scale = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
tscale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1., 's': 1e3}


def val(r, name):
  i = hdr.index(name)
  return float(r[i].replace(',', ''))*scale[units[i]]


caps = []
assert len(data) == len(EXPECT), (len(data), [r[hdr.index('Kernel Name')] for r in data])
for r, (kern, S, f16, alg) in zip(data, EXPECT):
  name = r[hdr.index('Kernel Name')]
  assert kern in name, (kern, name)
  ti = hdr.index('gpu__time_duration.sum')
  # the last template argument of the loop kernels is CH16 (> 0: float16 pre-filter)
  assert (not name.split('>')[0].rstrip().endswith(' 0')) == f16 or kern == 'scan_kernel', (name, f16)
  c = {'kernel': kern, 'S': S, 'filter16': f16, 'kernel_name': name, 'dram_bytes_read': val(r, 'dram__bytes_read.sum'),
       'dram_bytes_write': val(r, 'dram__bytes_write.sum'), 'algorithmic_bytes': alg,
       'duration_ms': float(r[ti].replace(',', ''))*tscale[units[ti]]}
  c['dram_bytes'] = c['dram_bytes_read'] + c['dram_bytes_write']
  c['ratio'] = c['dram_bytes']/alg
  c['dram_GBps_under_ncu'] = c['dram_bytes']/(c['duration_ms']*1e-3)/1e9
  caps.append(c)
json.dump({'captures': caps, 'command': 'ncu --set full --clock-control none -k regex:greedy_loop_kernel|omp_loop_kernel|scan_kernel python tools/ncu_traffic_case.py'},
          open(out_path, 'w'), indent=1)
print(json.dumps(caps, indent=1))
