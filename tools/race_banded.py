"""racecheck target: the two-phase banded path alone (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import uniform_blocks, vector
import qrkit_b200 as qk
os.environ["QRK_BANDED_GROUP"] = "3"
slabs = uniform_blocks(11, 16, 24)
s = qk.BandedBlockedSparseQR(slabs, num_blocks=11, block_rows=16, block_cols=24, overlap=16)
print(float(np.abs(s.solve(vector(11 * 16, seed=3))).max()))
