#!/usr/bin/env python3
"""bench.py — MAF columns/sec of the PhyloCSF++ hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path (pack -> site-pattern keys/dedup -> Felsenstein pruning of both ECMs on
every codon window of both strands -> deciban scatter, plus the per-base BLS) over one batch of synthetic
alignment columns of BASELINE.json's config 3 shape: model 58mammals, 30 % missing cells, chromosome-scale
chain cut into batches of --cols columns (100 M columns = 12 such batches; every batch is larger than L2).

  value  columns/s with the batch resident in HBM (device-pointer C-ABI, CUDA events, max over ranks)
  e2e    columns/s through the host-buffer C-ABI call (pcsf_tracks): pinned host input, H2D, kernels, D2H
  --impl reference   the reference's own CPU implementation (oracle/_ref/phylocsf_ref = its unmodified sources compiled
                     against the GSL shim, OpenMP over all host cores; the oracle port only if that binary is missing),
                     MAF file -> 7 wig files on a bounded sample of the same workload

Launch: python bench.py --gpus N --steps K --warmup W   (N > 1: under torch.distributed.run, one rank per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "MAF columns/sec (6 frames, coding+noncoding ECM)"
UNIT = "columns/s"


def flops_per_pruning(nl: int) -> int:
    """internal-child mat-vecs + Hadamard products + root dot."""
    return 2 * 64 * 64 * (nl - 2) + 64 * (nl - 1) + 128


def hbm_passes(tstats, nl, B, Wn):
    peak = 6458.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    alg = {"k_pack": ("ms_pack", 2 * nl * B),                       # ASCII in, codes out
           "k_keys_tracks": ("ms_hash", nl * B + 32 * B),           # codes in, two 128-bit keys per column out
           "dedup (insert/resolve/scan/finalize)": ("ms_dedup", 40 * 2 * Wn),
           "k_scatter_tracks": ("ms_scatter", 24 * 2 * Wn),
           "k_bls": ("ms_bls", nl * B + 8 * B)}
    out = {}
    for k, (key, nbytes) in alg.items():
        ms = tstats.get(key) or 0.0
        if ms > 0:
            gbs = nbytes / (ms * 1e-3) / 1e9
            out[k] = {"ms": ms, "bytes": nbytes, "gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    return out


# ------------------------------------------------------------------------------------------------ CPU arm
_W = {}


def _cpu_init(model_name):
    from oracle import oracle as orc
    from phylocsfpp_b200.models import load_model
    m = load_model(model_name)
    _W["m"] = m
    _W["mc"] = orc.OracleModel(m.tree, m.S_c, m.f_c)
    _W["mnc"] = orc.OracleModel(m.tree, m.S_nc, m.f_nc)


def _cpu_chunk(seqs):
    """Reference path for one slice of columns: both strands' codon windows, both models, BLS."""
    from oracle import oracle as orc
    plus, minus = orc.window_codons(seqs)
    a = orc.run_tracks(_W["mc"], _W["mnc"], plus)
    b = orc.run_tracks(_W["mc"], _W["mnc"], minus)
    c = orc.bls(_W["m"].tree, seqs)[1]
    return float(a.sum() + b.sum() + c.sum())


def cpu_columns_per_sec(model_name, seqs_host: np.ndarray, ncores: int, pool=None):
    """Times the CPU path over seqs_host [nl, S]; returns (columns/s, seconds)."""
    import multiprocessing as mp
    S = seqs_host.shape[1]
    per = max(256, S // (ncores * 4))
    chunks = [np.ascontiguousarray(seqs_host[:, i:min(S, i + per + 2)]) for i in range(0, S - 2, per)]
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(ncores, initializer=_cpu_init, initargs=(model_name,))
        pool.map(_cpu_chunk, chunks[:ncores])  # warm: model build in every worker
    t0 = time.perf_counter()
    pool.map(_cpu_chunk, chunks)
    dt = time.perf_counter() - t0
    if own:
        pool.close()
    return S / dt, dt


REF_BIN = os.path.join(ROOT, "oracle", "_ref", "phylocsf_ref")


def ref_sample_shape(ncores: int, per_step: bool):
    """(chains, columns per chain) of the CPU sample: chains are the reference's unit of parallel work (a reader job owns the
    chains that start in its byte range, parallel_file_reader.hpp:281-350), so the sample is cut into 4 (2) chains per core;
    3000-4000 columns per chain keep the per-chain model instantiation (instance.hpp:449-646) near 10 % as on real chains."""
    return (2 * ncores, 3000) if per_step else (4 * ncores, 4000)


class RefSample:
    """The first S columns of the workload written as a MAF file (tmpfs when available) for the reference binary."""

    def __init__(self, model, mat: np.ndarray, chains: int, chain_cols: int):
        import tempfile
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from make_synth_maf import write_synth_maf
        self.S = chains * chain_cols
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        self.dir = tempfile.mkdtemp(prefix="pcsf_ref_", dir=base)
        self.maf = os.path.join(self.dir, "sample.maf")
        self.info = write_synth_maf(self.maf, model, self.S, seed=7, mat=mat, chain_cols=chain_cols, hole_p=0.0, ref_gap=0.0, alien_p=0.0)

    def run(self, model_name: str, threads: int) -> float:
        """One build-tracks run of the reference (power + 6 raw tracks); returns seconds."""
        import shutil
        out = os.path.join(self.dir, "out")
        shutil.rmtree(out, ignore_errors=True)
        t0 = time.perf_counter()
        subprocess.run([REF_BIN, "build-tracks", "--threads", str(threads), "--output", out, model_name, self.maf], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dt = time.perf_counter() - t0
        assert os.path.getsize(os.path.join(out, "PhyloCSFRaw+1.wig")) > 0
        return dt

    def close(self):
        import shutil
        shutil.rmtree(self.dir, ignore_errors=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="58mammals")
    ap.add_argument("--cols", type=int, default=1 << 23, help="alignment columns per step (per GPU)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="columns of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-port", action="store_true", help="time the oracle port instead of the reference binary on the CPU legs")
    ap.add_argument("--no-dedup", action="store_true")
    ap.add_argument("--precision", default="tc5", choices=["f64", "tc5"],
                    help="tc5: tcgen05/TMEM split-TF32 path (fastest path inside the 1e-3 deciban contract); "
                         "f64: FP64 DMMA path (parity anchor, byte-identical wig text)")
    args = ap.parse_args()

    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version banner there at
    # communicator creation) are sent to stderr instead
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ncores = os.cpu_count() or 1

    from phylocsfpp_b200.models import load_model
    model = load_model(args.model)
    nl = model.nl
    workload = (f"build-tracks {args.model} on synthetic hg38.100way-shaped alignment columns, 30% missing cells, "
                f"batches of {args.cols} columns (100M-column chain = {-(-100_000_000 // args.cols)} batches)")

    if args.impl == "reference":
        # rank 0 only: the reference's own CPU implementation on a bounded sample per step
        if rank != 0:
            return
        import torch
        from phylocsfpp_b200.synth import synth_alignment
        if os.path.exists(REF_BIN) and not args.cpu_port:
            chains, chain_cols = ref_sample_shape(ncores, per_step=True)
            S = args.cpu_sample or chains * chain_cols
            chains = max(1, S // chain_cols)
            seqs = synth_alignment(model, chains * chain_cols, seed=1234, device="cpu")[:, :chains * chain_cols].numpy()
            smp = RefSample(model, seqs, chains, chain_cols)
            S = smp.S
            for _ in range(args.warmup):
                smp.run(args.model, ncores)
            t = sum(smp.run(args.model, ncores) for _ in range(args.steps))
            smp.close()
            kind = "reference"
            sample = (f"{S} columns per step of the same synthetic workload as a MAF file of {chains} chains x {chain_cols} columns "
                      f"({smp.info['blocks']} blocks, {smp.info['bytes']} bytes, tmpfs), the reference's unmodified build-tracks "
                      f"(oracle/_ref, GSL shim, -O3 -fopenmp) --threads {ncores}: MAF parse + 6 raw tracks + power track + wig text")
        else:
            import multiprocessing as mp
            S = args.cpu_sample or 6000 * ncores
            seqs = synth_alignment(model, S, seed=1234, device="cpu")[:, :S].numpy()
            pool = mp.get_context("fork").Pool(ncores, initializer=_cpu_init, initargs=(args.model,))
            pool.map(_cpu_chunk, [np.ascontiguousarray(seqs[:, :64])] * ncores)
            for _ in range(args.warmup):
                cpu_columns_per_sec(args.model, seqs, ncores, pool)
            t = 0.0
            for _ in range(args.steps):
                t += cpu_columns_per_sec(args.model, seqs, ncores, pool)[1]
            pool.close()
            kind = "port"
            sample = f"{S} columns per step of the same synthetic workload, oracle port of the reference path, {ncores} processes"
        v = S * args.steps / t
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "sample_columns_per_step": S, "phylogenetic_model": args.model},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}), file=json_out, flush=True)
        return

    import torch
    import torch.distributed as dist
    from phylocsfpp_b200 import capi
    from phylocsfpp_b200.synth import synth_alignment

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: phylocsfpp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    B = args.cols
    Wn = B - 2
    dm = capi.DeviceModel(model, local_rank)
    seqs = synth_alignment(model, B, seed=1234 + rank, device=dev)     # [nl, ld] resident in HBM
    ld = seqs.shape[1]
    plus = torch.empty(Wn, dtype=torch.float64, device=dev)
    minus = torch.empty(Wn, dtype=torch.float64, device=dev)
    bls = torch.empty(B, dtype=torch.float64, device=dev)
    flags = (capi.TRACKS_SCORES | capi.TRACKS_BLS | (capi.TRACKS_NO_DEDUP if args.no_dedup else 0)
             | {"f64": 0, "tc5": capi.TRACKS_TC5}[args.precision])
    stream = torch.cuda.current_stream()

    def step():
        dm.tracks_device(seqs.data_ptr(), B, ld, flags, plus.data_ptr(), minus.data_ptr(), bls.data_ptr(), 0, stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    stats = dm.tracks_device_finish(stream.cuda_stream)

    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    stats = dm.tracks_device_finish(stream.cuda_stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    t_ms = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * B * args.steps / (ms_max / 1e3)
    checksum = float(plus[::4097].sum().item() + minus[::4099].sum().item() + bls[::4111].sum().item())

    # ---- e2e through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region)
    h_seqs = torch.empty((nl, ld), dtype=torch.uint8, pin_memory=True)
    h_seqs.copy_(seqs)
    h_plus = torch.empty(Wn, dtype=torch.float64, pin_memory=True)
    h_minus = torch.empty(Wn, dtype=torch.float64, pin_memory=True)
    h_bls = torch.empty(B, dtype=torch.float64, pin_memory=True)
    lib = capi.load()
    st = capi.TracksStats()

    def e2e_step():
        capi._check(lib.pcsf_tracks(dm.h, h_seqs.data_ptr(), B, ld, flags, h_plus.data_ptr(), h_minus.data_ptr(),
                                    h_bls.data_ptr(), None, st))

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * B * n_e2e / float(t_e2e.item())
    assert abs(float(h_bls[::4111].sum()) - float(bls[::4111].sum().item())) < 1e-6

    # ---- per-kernel durations (CUDA events on the launching stream, one extra instrumented step)
    dm.set_timing(True)
    step()
    tstats = dm.tracks_device_finish(stream.cuda_stream)
    dm.set_timing(False)

    # one instrumented step of the other precisions, for the side-by-side the north star asks for
    other = {}
    if rank == 0:
        base = flags & ~capi.TRACKS_TC5
        dm.set_timing(True)
        for pname, pflag in (("f64", 0), ("tc5", capi.TRACKS_TC5)):
            if pname == args.precision:
                continue
            dm.tracks_device(seqs.data_ptr(), B, ld, base | pflag, plus.data_ptr(), minus.data_ptr(), bls.data_ptr(), 0, stream.cuda_stream)
            o = dm.tracks_device_finish(stream.cuda_stream)
            other[pname] = {"ms_prune_per_step": o["ms_prune"],
                            "columns_per_s_kernels_only": B / (1e-3 * sum(o[k] for k in ("ms_pack", "ms_hash", "ms_dedup", "ms_prune", "ms_scatter", "ms_bls")))}
        dm.set_timing(False)

    if rank == 0:
        F = flops_per_pruning(nl)
        n_prune_launch = max(1, tstats["n_chunks"])
        flop_per_launch = tstats["n_unique"] * 2 * F / n_prune_launch
        ms_per_launch = tstats["ms_prune"] / n_prune_launch
        achieved = flop_per_launch / (ms_per_launch * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "profiles", "peaks_fp64.json")))
        except Exception:
            pass
        traffic = None
        traffic_note = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_prune_traffic.json")))
            traffic = tj.get({"f64": "dram_bytes_per_launch", "f32": "dram_bytes_per_launch_f32", "tc5": "dram_bytes_per_launch_tc5"}[args.precision])
            cap_cols = tj.get("tc5_columns_per_captured_launch") if args.precision == "tc5" else None
            if traffic and cap_cols:
                # the ncu --set full capture ran on a smaller batch; DRAM traffic of this kernel (leaf codes in, log z out) is linear
                # in the columns of a launch
                traffic = traffic * (B / n_prune_launch) / cap_cols
                traffic_note = f"dram__bytes_read+write of one ncu --set full capture at {cap_cols} columns per launch, scaled to {B // n_prune_launch}"
        except Exception:
            pass
        if args.precision == "f64":
            kernel, dtype = "k_prune", "f64"
            peak = peaks.get("micro", {}).get("dmma_tflops", 37.16)
            peak_source = ("FP64 DMMA (mma.sync.m8n8k4.f64) register-loop peak measured on this pool "
                           "(profiles/peaks_fp64.json); MEASURED_PEAKS.json carries no FP64 figure")
            executed = achieved
        else:
            # executed tensor flops per pruning: every inner edge is a (window x 64 x 64) product done as 3 TF32 products
            # (hi*hi, hi*lo, lo*hi) on either tensor path; leaf edges are gathers (no tensor work) on both
            kernel, dtype = "k_prune_tc5", "tf32x3/f32"
            peak = peaks.get("tf32_tflops", 764.2)
            peak_source = ("dense TF32 tensor peak = cuBLAS TF32 GEMM 8192^3 measured on this pool (profiles/peaks_fp64.json; "
                           "MEASURED_PEAKS.json carries bf16 only: 1605 TFLOP/s burst, TF32 is half rate); `achieved` counts "
                           "ALGORITHMIC flops (one FP product per term), `executed_tflops` what the tensor pipe ran")
            mult = (3.0 * 8192 * (nl - 2)) / F
            executed = achieved * mult
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload, "phylogenetic_model": args.model, "leaves": nl, "columns_per_step_per_gpu": B,
                       "precision": args.precision,
                       "missing_fraction": 0.30, "l2_policy": "inputs larger than L2 (%.0f MB per step)" % (nl * B / 1e6),
                       "dedup": not args.no_dedup, "unique_pattern_ratio": tstats["n_unique"] / max(1, tstats["n_windows"]),
                       "sharding": "independent column batches per rank, no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nl * B, "d2h_bytes_per_step": 8 * (2 * Wn + B),
                    "steps": n_e2e, "api": "pcsf_tracks (C-ABI, pinned host buffers)"},
            "gpu_launches": stats["n_launches"] * args.steps,
            "clocks": clocks,
            "roofline": {"kernel": kernel, "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "executed_tflops": executed,
                         "executed_frac": executed / peak, "peak_source": peak_source,
                         "flop_per_launch": flop_per_launch, "ms_per_launch": ms_per_launch,
                         "prunings_per_launch": tstats["n_unique"] * 2 / n_prune_launch},
            "stages_ms": {k: tstats[k] for k in ("ms_pack", "ms_hash", "ms_dedup", "ms_prune", "ms_scatter", "ms_bls")},
            # the HBM-bound passes around the pruning kernel: ALGORITHMIC bytes / measured time, against
            # MEASURED_PEAKS.json's copy bandwidth
            "hbm_passes": hbm_passes(tstats, nl, B, Wn),
            "checksum": checksum,
            "other_precisions": other,
        }
        if world == 1 and not args.no_cpu_baseline:
            if os.path.exists(REF_BIN) and not args.cpu_port:
                chains, chain_cols = ref_sample_shape(ncores, per_step=False)
                if args.cpu_sample:
                    chains = max(1, args.cpu_sample // chain_cols)
                smp = RefSample(model, seqs[:, :chains * chain_cols].cpu().numpy(), chains, chain_cols)
                dt = smp.run(args.model, ncores)
                smp.close()
                out["cpu_baseline"] = {
                    "value": smp.S / dt, "unit": UNIT, "cores": ncores, "kind": "reference", "seconds": dt,
                    "sample": f"first {smp.S} columns of the step's batch as a MAF file of {chains} chains x {chain_cols} columns "
                              f"({smp.info['blocks']} blocks, tmpfs), the reference's unmodified build-tracks (oracle/_ref, GSL shim, "
                              f"-O3 -fopenmp) --threads {ncores}: MAF parse + 6 raw tracks + power track + wig text"}
            else:
                S = args.cpu_sample or 3000 * ncores
                v, dt = cpu_columns_per_sec(args.model, seqs[:, :S].cpu().numpy(), ncores)
                out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": ncores, "kind": "port", "seconds": dt,
                                       "sample": f"first {S} columns of the step's batch, oracle port of the reference path, "
                                                 f"{ncores} processes"}
        print(json.dumps(out), file=json_out, flush=True)
    dm.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
