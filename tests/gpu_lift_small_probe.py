"""lifting time at small sizes = cost of the lower levels of a big frame (development aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, dwt_b200 as D
from oracle import pyoracle as O
cod = D.Codec()
for (w, h) in [(120, 68), (240, 135), (480, 270), (960, 540), (1920, 1080), (3840, 2160), (7680, 4320)]:
    img = O.synth(w, h, 'photo', 1)
    fe, fd = [], []
    for rep in range(8):
        s = cod.encode(img); fe.append(cod.stats.ms_lift)
        d = cod.decode(s); fd.append(cod.stats.ms_lift)
    print(w, h, 'lift fwd min %.4f med %.4f  inv min %.4f med %.4f ms' %
          (min(fe), sorted(fe)[4], min(fd), sorted(fd)[4]), flush=True)
