"""encode + decode one synthetic frame (profiling aid); usage: [W H [reps [in_flight]]]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, dwt_b200 as D
from oracle import pyoracle as O
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cod = D.Codec()
if len(sys.argv) > 4:
    cod.set_in_flight(int(sys.argv[4]))  # >= 4: the throughput variant of the scan (dec_scan_serial_kernel)
img = O.synth(w, h, 'photo', 1)
s = cod.encode(img)
for _ in range(reps):
    d = cod.decode(s)
print('ok' if (d == img).all() else 'MISMATCH', cod.stats.ms_total, cod.stats.ms_coder)
