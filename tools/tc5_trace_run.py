import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phylocsfpp_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libphylocsf_b200_trace.so")
from phylocsfpp_b200.models import load_model
from tests.util import random_alignment
model = load_model("58mammals")
seqs = random_alignment(model.nl, 148 * 256 * 3 // 2 + 2, seed=1, gap=0.3, conserve=0.7)
dm = capi.DeviceModel(model)
r = dm.tracks(seqs, bls=False, tc5=True, dedup=False)
dm.close()
