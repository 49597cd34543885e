import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phylocsfpp_b200 import capi
from phylocsfpp_b200.models import load_model
from tests.util import random_alignment
for name, L in (("29mammals", 60000), ("12flies", 60000), ("58mammals", 40000), ("100vertebrates", 30000)):
    model = load_model(name)
    seqs = random_alignment(model.nl, L, seed=4242, gap=0.3, conserve=0.7)
    dm = capi.DeviceModel(model)
    f64 = dm.tracks(seqs, bls=False, dedup=False)
    W = L - 2
    runs = []
    for rep in range(4):
        t5 = dm.tracks(seqs, bls=False, tc5=True, dedup=False)
        v = np.empty(2 * W); v[0::2] = t5["plus"]; v[1::2] = t5["minus"]
        runs.append(v)
    ref = np.empty(2 * W); ref[0::2] = f64["plus"]; ref[1::2] = f64["minus"]
    for rep, v in enumerate(runs):
        err = np.abs(v - ref)
        bad = np.nonzero(~(err < 2e-4))[0]
        print(name, "run", rep, "bad(>2e-4)", bad.size, "max", np.nanmax(err), "identical to run0:", np.array_equal(v, runs[0]),
              "bad idx", bad[:8], "tile parity", (bad[:8] // 128) % 2, "err", err[bad[:8]])
    dm.close()
