"""score-msa workload for ncu captures of the MLE / OMEGA kernels: config-5 shaped alignments (29mammals reduced to 12 species, 30..600 columns).
usage: python tools/ncu_msa.py mle|omega <n alignments>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phylocsfpp_b200 import capi
from phylocsfpp_b200.models import load_model
from phylocsfpp_b200.synth import synth_alignment
what, n = sys.argv[1], int(sys.argv[2])
model = load_model("29mammals", "Human,Chimp,Mouse,Dog,Cow,Horse,Elephant,Armadillo,Rat,Rabbit,Cat,Megabat")
rng = np.random.default_rng(5)
lens = np.exp(rng.uniform(np.log(30), np.log(600), n)).astype(np.int64)
mat = synth_alignment(model, int(lens.sum()), seed=11, device="cuda")[:, :int(lens.sum())].cpu().numpy()
starts = np.cumsum(lens) - lens
alns = [np.ascontiguousarray(mat[:, s:s + l]) for s, l in zip(starts, lens)]
dm = capi.DeviceModel(model, 0)
p, a, b = dm.score_msa(alns, capi.STRATEGY_MLE if what == "mle" else capi.STRATEGY_OMEGA)
print(what, n, "finite", int(np.isfinite(p).sum()), dm.score_msa_stats() if what == "mle" else "")
dm.close()
