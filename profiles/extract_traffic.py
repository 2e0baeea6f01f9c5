#!/usr/bin/env python
"""profiles/extract_traffic.py <ncu-rep> <algorithmic bytes of the captured launch> <S>
Writes profiles/r01_loop_traffic.json: DRAM traffic of the captured greedy_loop_kernel launch."""
import csv
import json
import os
import subprocess
import sys

rep, alg_bytes, S = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2]
scale = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def val(name):
  i = hdr.index(name)
  return float(data[i].replace(',', '')) * scale[units[i]]


out = {'kernel': 'greedy_loop_kernel', 'S': S, 'kernel_name': data[hdr.index('Kernel Name')],
       'dram_bytes_read': val('dram__bytes_read.sum'), 'dram_bytes_write': val('dram__bytes_write.sum'),
       'algorithmic_bytes': alg_bytes, 'duration_ms': float(data[hdr.index('gpu__time_duration.sum')].replace(',', '')),
       'duration_unit': units[hdr.index('gpu__time_duration.sum')]}
out['dram_bytes'] = out['dram_bytes_read'] + out['dram_bytes_write']
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'r01_loop_traffic.json')
json.dump(out, open(path, 'w'), indent=1)
print(json.dumps(out))
