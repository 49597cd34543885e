"""Runs k_prune_tc5 from the PCSF_TC5_TRACE build of the library (per-step clock64 timestamps of one CTA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PCSF_TC5_NIDS", "1")   # the trace buffer is 12 KB of static shared memory
import numpy as np
import torch
from phylocsfpp_b200 import capi
capi.LIB_PATH = os.path.join(os.path.dirname(capi.LIB_PATH), "libphylocsf_b200_trace.so")
from phylocsfpp_b200.models import load_model
from phylocsfpp_b200.synth import synth_alignment
model = load_model("58mammals")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 6 + 2
seqs = synth_alignment(model, B, seed=1234, device="cpu")[:, :B].numpy()
dm = capi.DeviceModel(model)
r = dm.tracks(seqs, bls=False, tc5=True, dedup=False)
dm.close()
