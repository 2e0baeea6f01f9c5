#!/usr/bin/env python
"""profiles/extract_traffic.py <ncu-rep> <out.json>
DRAM traffic (ncu dram__bytes_read.sum + dram__bytes_write.sum) of the captured greedy_loop_kernel / scan_kernel launches
of tools/ncu_traffic_case.py next to their algorithmic bytes (4 N S per pass).  bench.py reads the ratio from it
(roofline.traffic)."""
import csv
import json
import subprocess
import sys

rep, out_path = sys.argv[1], sys.argv[2]
# (kernel substring, S, algorithmic bytes per launch) in launch order of tools/ncu_traffic_case.py
EXPECT = [('greedy_loop_kernel', 512, 4.*10_000_000*512*5), ('omp_loop_kernel', 512, 4.*10_000_000*512*2), ('scan_kernel', 512, 4.*10_000_000*512),
          ('scan_kernel', 512, 4.*10_000_000*512), ('greedy_loop_kernel', 256, 4.*1_000_000*256*5)]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if 'exact_scan_kernel' not in r[rows[0].index('Kernel Name')]]
scale = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}
tscale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1., 's': 1e3}


def val(r, name):
  i = hdr.index(name)
  return float(r[i].replace(',', ''))*scale[units[i]]


caps = []
for r, (kern, S, alg) in zip(data, EXPECT):
  name = r[hdr.index('Kernel Name')]
  assert kern in name, (kern, name)
  ti = hdr.index('gpu__time_duration.sum')
  c = {'kernel': kern, 'S': S, 'kernel_name': name, 'dram_bytes_read': val(r, 'dram__bytes_read.sum'),
       'dram_bytes_write': val(r, 'dram__bytes_write.sum'), 'algorithmic_bytes': alg,
       'duration_ms': float(r[ti].replace(',', ''))*tscale[units[ti]]}
  c['dram_bytes'] = c['dram_bytes_read'] + c['dram_bytes_write']
  c['ratio'] = c['dram_bytes']/alg
  c['dram_GBps_under_ncu'] = c['dram_bytes']/(c['duration_ms']*1e-3)/1e9
  caps.append(c)
json.dump({'captures': caps, 'command': 'ncu --set full --clock-control none -k regex:greedy_loop_kernel|omp_loop_kernel|scan_kernel python tools/ncu_traffic_case.py'},
          open(out_path, 'w'), indent=1)
print(json.dumps(caps, indent=1))
