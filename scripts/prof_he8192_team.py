import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from chord_detection_b200 import ops, synth
dev = torch.device("cuda:0")
seg = torch.from_numpy(synth.s_poly_long(5, 22050, 1 << 22)).to(dev)
x = seg.repeat(44)[: 22050 * 8192].contiguous()
os.environ["CDB_HE8192"] = "team"
for _ in range(3):
    ops.harmonic_energy(x, 22050)
torch.cuda.synchronize()
