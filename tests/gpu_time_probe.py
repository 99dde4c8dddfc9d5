"""encode/decode timing probe on synthetic images (development aid)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, dwt_b200 as D
from oracle import pyoracle as O
cod = D.Codec()
cases = [(1920, 1080, 'photo'), (3840, 2160, 'photo'), (7680, 4320, 'photo'), (7680, 4320, 'noise')]
if len(sys.argv) > 1:
    cases = [(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])]
for (w, h, kind) in cases:
    img = O.synth(w, h, kind, 1)
    for rep in range(2):
        t = time.time(); s = cod.encode(img); te = time.time() - t
    st = cod.stats
    print(w, h, kind, 'bytes', len(s), 'enc wall %.1f ms dev total %.2f lift %.3f lin %.3f coder %.3f' %
          (te * 1e3, st.ms_total, st.ms_lift, st.ms_linearize, st.ms_coder), flush=True)
    for rep in range(2):
        t = time.time(); d = cod.decode(s); dt = time.time() - t
    st = cod.stats
    print('   dec', 'ok' if (d == img).all() else 'MISMATCH', 'wall %.1f ms' % (dt * 1e3),
          'dev total %.2f coder %.2f recon %.2f lift %.2f' % (st.ms_total, st.ms_coder, st.ms_linearize, st.ms_lift),
          'windows', st.parse_windows, 'jumps', st.parse_jumps, 'exact', st.parse_exact, flush=True)
