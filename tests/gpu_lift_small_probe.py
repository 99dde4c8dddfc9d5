"""lifting time per image size (development aid); usage: [W H [reps]]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, dwt_b200 as D
from oracle import pyoracle as O
cod = D.Codec()
sizes = [(120, 68), (240, 135), (480, 270), (960, 540), (1920, 1080), (3840, 2160), (7680, 4320)]
reps = 8
if len(sys.argv) > 2:
    sizes = [(int(sys.argv[1]), int(sys.argv[2]))]
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
for (w, h) in sizes:
    img = O.synth(w, h, 'photo', 1)
    fe, fd = [], []
    for rep in range(reps):
        s = cod.encode(img); fe.append(cod.stats.ms_lift)
        d = cod.decode(s); fd.append(cod.stats.ms_lift)
    print(w, h, 'lift fwd min %.4f med %.4f  inv min %.4f med %.4f ms' %
          (min(fe), sorted(fe)[reps // 2], min(fd), sorted(fd)[reps // 2]), flush=True)
