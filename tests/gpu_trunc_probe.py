import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, dwt_b200 as D
from oracle import pyoracle as O
cod = D.Codec()
img = O.synth(7680, 4320, 'photo', 1)
s = cod.encode(img, 1048576)
for _ in range(3):
    d = cod.decode(s)
print(d.shape, cod.stats.ms_total, cod.stats.ms_coder)
