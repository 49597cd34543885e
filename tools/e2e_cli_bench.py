#!/usr/bin/env python3
"""End-to-end timings of the command line host on synthetic MAF files (BASELINE configs 3 and 5 shapes, scaled down):
wall clock of `phylocsf_b200 build-tracks` / `score-msa` from a page-cached MAF file to the output files.
usage: tools/e2e_cli_bench.py [tracks_columns] [msa_columns] [gpus]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
BIN = os.path.join(ROOT, "phylocsfpp_b200", "bin", "phylocsf_b200")


def run(cmd):
    t0 = time.perf_counter()
    out = subprocess.run(cmd, check=True, capture_output=True, text=True, env=dict(os.environ, PCSF_HOST_STATS="1")).stdout
    dt = time.perf_counter() - t0
    stats = [ln for ln in out.splitlines() if ln.startswith("{")]
    return dt, (json.loads(stats[-1]) if stats else None)


def main():
    from make_synth_maf import write_synth_maf
    from phylocsfpp_b200.models import load_model
    n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    n_msa = int(sys.argv[2]) if len(sys.argv) > 2 else 6_000_000
    gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    threads = os.cpu_count() or 1
    res = {"host_threads": threads, "gpus": gpus}
    with tempfile.TemporaryDirectory() as tmp:
        if n_tracks:
            maf = os.path.join(tmp, "tracks.maf")
            info = write_synth_maf(maf, load_model("58mammals"), n_tracks, seed=7)
            open(maf, "rb").read()
            for prec in ("tc5", "f64"):
                best = None
                for rep in range(2):
                    dt, st = run([BIN, "build-tracks", "--threads", str(threads), "--gpus", str(gpus), "--precision", prec, "--output",
                                  os.path.join(tmp, "o_" + prec), "58mammals", maf])
                    if best is None or dt < best[0]:
                        best = (dt, st)
                res["build_tracks_" + prec] = {"maf": info, "process_seconds": best[0], "columns_per_s_process": info["columns"] / best[0], "tool_stats": best[1]}
        if n_msa:
            maf = os.path.join(tmp, "blocks.maf")
            info = write_synth_maf(maf, load_model("29mammals"), n_msa, seed=3, loguniform_blocks=(30, 600))
            open(maf, "rb").read()
            species = "Human,Chimp,Mouse,Dog,Cow,Horse,Elephant,Armadillo,Rat,Rabbit,Cat,Megabat"
            for strat in ("fixed", "mle"):
                dt, st = run([BIN, "score-msa", "--strategy", strat, "--comp-anc", "1", "--gpus", str(gpus), "--species", species, "--output",
                              os.path.join(tmp, "s_" + strat), "29mammals", maf])
                res["score_msa_" + strat] = {"maf": info, "process_seconds": dt, "alignments_per_s": info["blocks"] / dt,
                                             "columns_per_s": info["columns"] / dt, "tool_stats": st}
            # OMEGA: ~200 likelihood evaluations and ~100 eigendecompositions per alignment -> a tenth of the file
            maf_o = os.path.join(tmp, "blocks_omega.maf")
            info_o = write_synth_maf(maf_o, load_model("29mammals"), max(30000, n_msa // 10), seed=4, loguniform_blocks=(30, 600))
            dt, st = run([BIN, "score-msa", "--strategy", "omega", "--gpus", str(gpus), "--species", species, "--output", os.path.join(tmp, "s_omega"),
                         "29mammals", maf_o])
            res["score_msa_omega"] = {"maf": info_o, "process_seconds": dt, "alignments_per_s": info_o["blocks"] / dt, "columns_per_s": info_o["columns"] / dt,
                                      "tool_stats": st}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
