"""GPU box: one synthetic 58mammals chromosome MAF (tmpfs), then build-tracks --precision tc5 under a list of host settings.
usage: python tools/cli_sweep.py <columns> <out.json> [NAME=ENV1=v,ENV2=v ...]"""
import json, os, subprocess, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from phylocsfpp_b200.models import load_model
from phylocsfpp_b200.synth import synth_alignment

cols, outp = int(sys.argv[1]), sys.argv[2]
settings = [("default", {})]
for a in sys.argv[3:]:
    name, envs = a.split("=", 1)
    settings.append((name, dict(kv.split("=") for kv in envs.split(","))))
BIN = os.path.join(ROOT, "phylocsfpp_b200", "bin", "phylocsf_b200")
model = load_model("58mammals")
tmp = tempfile.mkdtemp(prefix="pcsf_sweep_", dir="/dev/shm")
res = {}
try:
    mm = np.lib.format.open_memmap(os.path.join(tmp, "m.npy"), mode="w+", dtype=np.uint8, shape=(model.nl, cols))
    piece = 1 << 23
    for c0 in range(0, cols, piece):
        n = min(piece, cols - c0)
        mm[:, c0:c0 + n] = synth_alignment(model, n, seed=5000 + c0 // piece, device="cuda")[:, :n].cpu().numpy()
    off = mm.offset
    del mm
    maf = os.path.join(tmp, "chr1.maf")
    subprocess.run([BIN, "matrix-to-maf", "--chain", "25000000", "--skip-bytes", str(off), "58mammals", os.path.join(tmp, "m.npy"), str(cols), maf], check=True, capture_output=True)
    os.unlink(os.path.join(tmp, "m.npy"))
    torch.cuda.empty_cache()
    for name, env in settings:
        runs = []
        for rep in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([BIN, "build-tracks", "--threads", str(os.cpu_count()), "--precision", "tc5", "--output", os.path.join(tmp, "out"), "58mammals", maf],
                               check=True, capture_output=True, text=True, env=dict(os.environ, PCSF_HOST_STATS="1", **env))
            dt = time.perf_counter() - t0
            js = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
            runs.append({"process_seconds": dt, "stats": js})
        res[name] = runs
        print(name, json.dumps(runs[-1]), flush=True)
finally:
    shutil.rmtree(tmp, ignore_errors=True)
json.dump(res, open(outp, "w"), indent=1)
