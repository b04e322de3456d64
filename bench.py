#!/usr/bin/env python
"""Benchmark of the FAKEBOB NES attack hot path on B200.

  python bench.py [--config C2|C3|C4|C5] --gpus N --steps K --warmup W      (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference [--config ...] --gpus N --steps K --warmup W

Workloads (BASELINE.json `configs`, made concrete in BASELINE.md section 3; data and models are synthetic, seeded):
  C2 (default, the metric's config)  gmm_OSI untargeted, UBM + 5 MAP-adapted speaker GMMs x 2048 mixtures, S = 50
  C3  iv_SV, 2048-mixture full UBM, 400-dim i-vector, LDA 200 + PLDA, S = 50
  C4  gmm_CSI targeted, 5 speaker GMMs (no UBM), S = 256, antithetic pairs sharded over the ranks (one all-reduce / iteration)
  C5  iv_OSI, 10 enrolled speakers, S = 512, 32 concurrent attack utterances sharded over the ranks (no collective)
One "step" = one NES iteration (FAKEBOB.py:168-214) of every utterance of the workload: S + 1 audios perturbed, quantised,
MFCC'd, VAD'd, delta/CMN'd, scored, loss, gradient estimate, momentum / sign / clip update.  Early stop is disabled
(unreachable threshold) so every iteration does the full work.

Prints ONE JSON line (rank 0).  DESIGN.md section 5 explains every key.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 80000
N_MIX = 2048
EPSILON = 0.002
THETA = 1e6                                           # unreachable: early stop never fires
IV_R, IV_L = 400, 200

CONFIGS = {
    "C2": dict(arch="gmm", task="OSI", attack="untargeted", K=5, S=50, utts=1, shard="pairs",
               desc="C2: gmm_OSI untargeted, UBM+5 spk x 2048 mix, samples_per_draw=50, 5 s @ 16 kHz, eps=0.002"),
    "C3": dict(arch="iv", task="SV", attack="untargeted", K=1, S=50, utts=1, shard="pairs",
               desc="C3: iv_SV, 2048-mix full UBM, 400-dim i-vector, LDA 200 + PLDA, samples_per_draw=50, 5 s @ 16 kHz"),
    "C4": dict(arch="gmm", task="CSI", attack="targeted", K=5, S=256, utts=1, shard="pairs",
               desc="C4: gmm_CSI targeted, 5 spk x 2048 mix (no UBM), samples_per_draw=256 sharded over the ranks, 5 s @ 16 kHz"),
    "C5": dict(arch="iv", task="OSI", attack="untargeted", K=10, S=512, utts=32, shard="utterances",
               desc="C5: iv_OSI 10-speaker open set, samples_per_draw=512, 32 concurrent 5 s utterances sharded over the ranks"),
}


def config_dict(name):
    """Identical in both arms (the driver compares them)."""
    c = CONFIGS[name]
    return {"workload": c["desc"], "name": name, "samples_per_draw": c["S"], "utterances": c["utts"], "speakers": c["K"],
            "mixtures": N_MIX, "audio_samples": N_SAMPLES}


def env_int(name, default):
    return int(os.environ.get(name, default))


# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons during the timed region (B200_PROFILING.md recipe), sampled through NVML
    every 20 ms (falls back to the nvidia-smi query line when pynvml is unavailable).  The sampler is stopped before the
    host-timed end-to-end legs: NVML queries take a driver lock that kernel launches also need."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []          # (sm_mhz, max_mhz, power_w, hw_slowdown, hw_thermal, sw_thermal, sw_power_cap)
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        hw = bool(r & 0x8)            # nvmlClocksThrottleReasonHwSlowdown
        hwt = bool(r & 0x40)          # HwThermalSlowdown
        swt = bool(r & 0x20)          # SwThermalSlowdown
        swp = bool(r & 0x4)           # SwPowerCap
        self.samples.append((sm, self.max_mhz, pw, hw, hwt, swt, swp))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        p = [x.strip() for x in out.strip().split(",")]
        if len(p) >= 7:
            act = [x.lower().startswith("active") for x in p[3:7]]
            self.samples.append((float(p[0]), float(p[1]), float(p[2]), act[0], act[1], act[2], act[3]))

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            time.sleep(0.02 if self.nvml is not None else 0.15)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [name for i, name in enumerate(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"))
                   if any(s[3 + i] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.samples[0][1], "reasons": reasons,
                "power_w_max": max(s[2] for s in self.samples), "samples": len(self.samples),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# -------------------------------------------------------------------------------------------------- workloads
def build_gmm_tree_gpu(root, device, n_speakers):
    """Synthetic pre-models/ + model/ tree whose features come from the CUDA front-end."""
    from fakebob_b200 import synth
    from fakebob_b200.engine import GmmEngine
    r = np.random.default_rng(0)
    dummy = {"weights": np.full(128, 1 / 128, np.float32), "means_invvars": r.standard_normal((128, 72)).astype(np.float32),
             "inv_vars": np.ones((128, 72), np.float32), "gconsts": np.zeros(128, np.float32)}
    fe = GmmEngine([dummy], device=device)
    tree = synth.build_gmm_tree(root, fe.features, n_speakers=n_speakers, C=N_MIX, n_ubm_utts=64, n_samples=N_SAMPLES)
    fe.close()
    return tree


def build_gmm_tree_cpu(root, n_speakers):
    from fakebob_b200 import synth
    from oracle import kaldi_feats as kf
    return synth.build_gmm_tree(root, kf.voiced_features, n_speakers=n_speakers, C=N_MIX, n_ubm_utts=64, n_samples=N_SAMPLES)


def build_iv_speakers_gpu(root, tree, device, n_speakers):
    """final.ubm / final.ie / back-end + enrolled speakers (*.iv 5-lists) with i-vectors extracted on the device."""
    from fakebob_b200 import synth
    from fakebob_b200.engine import IvectorEngine
    synth.build_ivector_params(root, tree["ubm_params"], R=IV_R, L=IV_L)
    eng = IvectorEngine(tree["pre_model_dir"], device=device)

    def ivec(w):
        return eng.extract_ivectors([np.ascontiguousarray(w, dtype=np.int16)])[0]

    def plda(enrolled, test):
        # the z-norm cohort of the synthetic workload only needs finite, speaker-dependent statistics (timing does not
        # depend on them); tests/test_gpu_fullsize.py checks real PLDA scores against the oracle fixture
        return np.arange(len(test), dtype=np.float64)[:, None] * 0.5 + 10.0 * np.arange(len(enrolled))[None, :]
    spk = synth.build_ivector_speakers(root, ivec, plda, n_speakers=n_speakers, n_samples=N_SAMPLES, n_znorm_utts=2)
    eng.close()
    return spk


def make_model(name, tree, iv_models, device, group_root):
    c = CONFIGS[name]
    if c["arch"] == "gmm":
        from fakebob_b200.gmm_ubm_CSI import gmm_CSI
        from fakebob_b200.gmm_ubm_OSI import gmm_OSI
        if c["task"] == "OSI":
            return gmm_OSI(os.path.join(group_root, "gmm-OSI"), tree["models"], tree["ubm"], pre_model_dir=tree["pre_model_dir"],
                           threshold=0.0, device=device)
        return gmm_CSI(os.path.join(group_root, "gmm-CSI"), tree["models"], pre_model_dir=tree["pre_model_dir"], device=device)
    from fakebob_b200.ivector_PLDA_OSI import iv_OSI
    from fakebob_b200.ivector_PLDA_SV import iv_SV
    if c["task"] == "SV":
        return iv_SV(os.path.join(group_root, "iv-SV"), iv_models[0], pre_model_dir=tree["pre_model_dir"], device=device)
    return iv_OSI(os.path.join(group_root, "iv-OSI"), iv_models, pre_model_dir=tree["pre_model_dir"], threshold=0.0, device=device)


def attack_kwargs(name):
    c = CONFIGS[name]
    kw = {"threshold": THETA}
    if c["task"] == "CSI":
        kw["target"] = 1
    return kw


def make_attacker(name, model, max_iter, seed):
    from fakebob_b200.FAKEBOB import FakeBob
    c = CONFIGS[name]
    # CSI has no threshold to make unreachable: a large margin kappa keeps the loss positive instead
    kappa = 1e3 if c["task"] == "CSI" else 0.0
    fb = FakeBob(c["task"], c["attack"], model, adver_thresh=kappa, epsilon=EPSILON, max_iter=max_iter,
                 samples_per_draw=c["S"], seed=seed, verbose=False)
    kw = attack_kwargs(name)
    fb.threshold = kw["threshold"]
    fb.target = kw.get("target")
    fb.true = None
    return fb


def utterances(name):
    from fakebob_b200 import synth
    return [synth.synth_utterance(u, u % 3, N_SAMPLES) for u in range(CONFIGS[name]["utts"])]


# -------------------------------------------------------------------------------------------------- CPU arm
def load_reference_fakebob():
    """The reference's own, unmodified FAKEBOB.py, staged by oracle/stage_reference.py (run by __graft_entry__.build()
    in the authoring container) into baseline/_ref -- /root/reference does not exist on the GPU box.  None if absent."""
    path = os.path.join(ROOT, "baseline", "_ref", "FAKEBOB.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_FAKEBOB", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.FakeBob


def oracle_model(name, tree, iv_root=None):
    from oracle import kaldi_files as kaldi_io           # the oracle's own model reader
    from oracle.diag_gmm import DiagGmm
    from oracle.scorers import OracleGmmCSI, OracleGmmOSI, OracleIvOSI, OracleIvSV
    c = CONFIGS[name]

    def lg(p):
        g = kaldi_io.read_diag_gmm(p)
        return DiagGmm(g["weights"], g["means_invvars"], g["inv_vars"], g["gconsts"])
    if c["arch"] == "gmm":
        if c["task"] == "OSI":
            return OracleGmmOSI(lg(tree["ubm"]), [lg(m[2]) for m in tree["models"]])
        return OracleGmmCSI([lg(m[2]) for m in tree["models"]], [m[3] for m in tree["models"]], [m[4] for m in tree["models"]])
    from fakebob_b200 import synth
    from oracle.ivector import load_system
    synth.build_ivector_params(iv_root, tree["ubm_params"], R=IV_R, L=IV_L)
    system = load_system(tree["pre_model_dir"])
    r = np.random.default_rng(5)
    enrolled = r.standard_normal((c["K"], IV_R))           # timing only: the enrolled identities do not change the work
    if c["task"] == "SV":
        return OracleIvSV(system, enrolled[:1], 0.0, 1.0)
    return OracleIvOSI(system, enrolled, np.zeros(c["K"]), np.ones(c["K"]))


def time_cpu(name, tree, iv_root, steps, warmup, budget_s=150.0):
    """Times the CPU path on the host cores: the reference's own FAKEBOB.py loop when staged (else its restatement
    oracle/nes.py) driving the oracle's numpy restatement of the Kaldi scoring.  Returns (iters_per_sec, sample, cores, loop)."""
    from oracle.nes import OracleFakeBob, PhiloxNoise
    c = CONFIGS[name]
    model = oracle_model(name, tree, iv_root)
    cores = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    RefFakeBob = load_reference_fakebob()
    loop = "reference FAKEBOB.py (unmodified, numpy RNG)" if RefFakeBob else "oracle/nes.py restatement of FAKEBOB.py (Philox noise on the host)"
    from fakebob_b200 import synth
    audio = synth.synth_utterance(0, 0, N_SAMPLES)
    kappa = 1e3 if c["task"] == "CSI" else 0.0

    def run(n_iter, s_draw):
        kw = dict(adver_thresh=kappa, epsilon=EPSILON, max_iter=n_iter, samples_per_draw=s_draw)
        akw = attack_kwargs(name)
        t0 = time.perf_counter()
        if RefFakeBob:
            import contextlib
            import io
            fb = RefFakeBob(c["task"], c["attack"], model, **kw)
            with contextlib.redirect_stdout(io.StringIO()):
                fb.attack(audio[:, None], os.devnull, n_jobs=1, **akw)
        else:
            fb = OracleFakeBob(c["task"], c["attack"], model, noise_fn=PhiloxNoise(1), **kw)
            fb.attack(audio, None, **akw)
        return (time.perf_counter() - t0) / n_iter
    # bounded sample: find a samples_per_draw whose (steps + warmup) iterations fit the budget, scale by audios scored
    s_draw = c["S"]
    t_first = run(1, min(s_draw, 4)) * (s_draw + 1) / (min(s_draw, 4) + 1)      # estimate of a full iteration
    while s_draw > 4 and (steps + warmup) * t_first * (s_draw + 1) / (c["S"] + 1) > budget_s:
        s_draw = max(4, s_draw // 2 - (s_draw // 2) % 2)
    scale = (s_draw + 1) / float(c["S"] + 1)
    if warmup > 1:
        run(warmup - 1, s_draw)
    per_iter = run(steps, s_draw)
    note = "%d x NES iteration of one utterance with samples_per_draw=%d (%d audios)" % (steps, s_draw, s_draw + 1)
    if s_draw != c["S"]:
        note += ", rate scaled by %d/%d to samples_per_draw=%d" % (s_draw + 1, c["S"] + 1, c["S"])
    if c["utts"] > 1:
        note += "; a step of the workload is %d such utterance-iterations (rate divided by %d)" % (c["utts"], c["utts"])
    return scale / per_iter / c["utts"], note, cores, loop


def run_reference(args, rank):
    if rank != 0:
        return
    name = args.config
    c = CONFIGS[name]
    root = tempfile.mkdtemp(prefix="fakebob_ref_")
    tree = build_gmm_tree_cpu(root, c["K"] if c["arch"] == "gmm" else 1)
    t0 = time.perf_counter()
    ips, sample, cores, loop = time_cpu(name, tree, root, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "NES attack iters/sec", "value": ips, "unit": "iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / ips, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(name),
        "note": "CPU path: %s driving the oracle (numpy restatement of the Kaldi arithmetic, BLAS on all host cores). The real "
                "Kaldi-subprocess path cannot run offline (no Kaldi, no pre-models); this baseline omits the reference's fork/exec "
                "and wav/ark/text I/O per score() call, so it is faster than the real reference (README.md:112-113: ~5 s / "
                "iteration GMM-UBM, ~12 s i-vector)." % loop,
        "cpu_baseline": {"value": ips, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample, "loop": loop},
        "e2e": {"value": ips, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------- roofline bookkeeping
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def load_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of the CURRENT kernels
    (profiles/ncu_traffic.json, written from the capture named inside it); None when the kernel has not been captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    e = t.get(kernel)
    return (e["dram_bytes_per_launch"], e["capture"]) if e else (None, None)


FP64_TFMA_PEAK = 17.1     # T FMA/s, DFMA, measured on this pool's B200 (DMMA m8n8k4: 18.6)


def rooflines(name, prof, rows, n_models, peaks, active_frac=None, b_local=None):
    """Per-kernel algorithmic work / measured stage time -> fraction of the governing measured peak.
    `rows` = voiced rows this rank scores per iteration.  The tensor-pipe kernels are timed inside a millisecond-scale window,
    so the burst cuBLAS figure is the denominator; the sustained one is given alongside."""
    c = CONFIGS[name]
    B = b_local if b_local else c["S"] + 1
    ms = {k: v[0] / max(v[1], 1) for k, v in prof.items() if v[1]}
    out = {}
    tf, tf_s, hbm = peaks["bf16_tflops"], peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]

    def tensor(stage, flops, kernel, note):
        if stage in ms and ms[stage] > 0:
            a = flops / (ms[stage] * 1e-3) / 1e12
            out[stage] = {"kernel": kernel, "bound": "tensor", "achieved": a, "peak": tf, "unit": "TFLOP/s", "frac": a / tf,
                          "frac_of_sustained_peak": a / tf_s, "algorithmic_flops_per_launch": flops, "kernel_ms": ms[stage], "note": note}

    def mem(stage, nbytes, kernel, note):
        if stage in ms and ms[stage] > 0:
            a = nbytes / (ms[stage] * 1e-3) / 1e9
            out[stage] = {"kernel": kernel, "bound": "hbm", "achieved": a, "peak": hbm, "unit": "GB/s", "frac": a / hbm,
                          "algorithmic_bytes_per_launch": nbytes, "kernel_ms": ms[stage], "note": note}
    if c["arch"] == "gmm":
        tensor("gmm", 2.0 * rows * N_MIX * 144 * n_models, "gmm_umma_kernel",
               "algorithmic = 2*rows*C*(2D) per model (Kaldi's two sgemv per frame, SURVEY 8d)")
    else:
        tensor("gmm", 2.0 * rows * N_MIX * 144, "gmm_umma_kernel<STORE>", "Gaussian selection on the diagonalised full UBM: 2*rows*C*(2D)")
        af = active_frac if active_frac else 1.0
        mem("ivec_lin", 4.0 * N_MIX * 72 * IV_R * af, "ivec_lin_tma_kernel",
            "bytes = active components x 72 x R x 4 (fp32 Sigma^-1 M rows of the components with gamma != 0; %.0f %% active)" % (100 * af))
        mem("ivec_quad", 4.0 * N_MIX * (IV_R * (IV_R + 1) // 2) * af, "ivec_quad_tma_kernel",
            "bytes = active components x R(R+1)/2 x 4 (fp32 U rows of the components with gamma != 0; %.0f %% active)" % (100 * af))
        tensor("fgmm_post", 2.0 * rows * 20 * (72 * 73 // 2 + 72), "fgmm_post_group_kernel",
               "20 full-covariance log-likelihoods per frame, 2*rows*20*(D(D+1)/2+D) FLOP; CUDA-core kernel, tensor peak shown for scale")
        mem("ivec_solve", 8.0 * B * (IV_R * (IV_R + 1) // 2) + 8.0 * B * IV_R * 2, "ivec_solve_kernel",
            "latency-bound blocked Cholesky in float64 (B*R^3/3 = %.2f GFLOP); bytes = packed posterior precision read + lin + solution" % (B * IV_R ** 3 / 3e9))
        mem("gselect", 4.0 * rows * N_MIX, "gselect_kernel", "bytes = component log-likelihoods read (rows x C x 4)")
        # the three float64 kernels against the measured FP64 FMA rate (scripts/fp64_probe.cu, profiles/r02_fp64_probe.txt)
        n_act = N_MIX * af
        for stage, fmas in (("ivec_lin", n_act * 72 * IV_R * B), ("ivec_quad", n_act * (IV_R * (IV_R + 1) // 2) * B),
                            ("ivec_solve", B * IV_R ** 3 / 6.0)):
            if stage in out:
                a = fmas / (ms[stage] * 1e-3) / 1e12
                out[stage]["fp64"] = {"achieved_tfma_per_s": a, "peak_tfma_per_s": FP64_TFMA_PEAK, "frac": a / FP64_TFMA_PEAK,
                                      "algorithmic_fma_per_launch": fmas}
    mem("mfcc", 2.0 * B * N_SAMPLES + 4.0 * B * (N_SAMPLES // 160) * 24, "mfcc_kernel", "bytes = int16 wave read + MFCC written (issue-bound kernel)")
    mem("feats", 4.0 * B * (N_SAMPLES // 160) * 24 + 2.0 * rows * 160 * 2, "feats_kernel", "bytes = MFCC read + fp16 hi/lo operand image written")
    return out, ms


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=os.environ.get("FB_BENCH_CONFIG", "C2"), choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the informational C3 leg of the default C2 run")
    ap.add_argument("--no-flush", action="store_true", help="skip the L2 flush between timed iterations (diagnostic)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    multi = world > 1
    if multi:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if multi:
            dist.barrier()

    def max_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    stream = torch.cuda.Stream()
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        if not args.no_flush:
            flush_buf.zero_()

    # ---- workload: rank 0 builds the synthetic tree(s), the others load them
    def build_trees(names):
        holder = [None]
        if rank == 0:
            root = tempfile.mkdtemp(prefix="fakebob_bench_")
            n_spk = max([CONFIGS[n]["K"] for n in names if CONFIGS[n]["arch"] == "gmm"] + [1])
            tree = build_gmm_tree_gpu(root, local_rank, n_spk)
            out = {k: tree[k] for k in ("pre_model_dir", "model_dir", "ubm", "spk_ids", "models")}
            out["root"] = root
            n_iv = max([CONFIGS[n]["K"] for n in names if CONFIGS[n]["arch"] == "iv"] + [0])
            if n_iv:
                out["iv_models"] = build_iv_speakers_gpu(root, tree, local_rank, n_iv)["models"]
            holder[0] = out
            ubm_params = tree["ubm_params"]
        if multi:
            dist.broadcast_object_list(holder, src=0)
        t = holder[0]
        if rank == 0:
            t["ubm_params"] = ubm_params
        return t

    def measure(name, tree, K, W, with_e2e=True):
        """All device / host timings of one configuration on this rank set."""
        c = CONFIGS[name]
        shard_utts = c["shard"] == "utterances"
        iv_models = tree.get("iv_models")
        model = make_model(name, tree, iv_models[:c["K"]] if iv_models else None, local_rank,
                           tempfile.mkdtemp(prefix="fakebob_grp_"))
        eng = model._engine
        eng.set_stream(stream.cuda_stream)
        if multi and not shard_utts:
            eng.comm_init_from_torch()
        utts = utterances(name)
        if shard_utts:
            lo, hi = (len(utts) * rank) // world, (len(utts) * (rank + 1)) // world
            mine = list(range(lo, hi))
        else:
            mine = [0]
        res = {"ms_flushed": 0.0, "ms_hot": 0.0, "launches": 0, "e2e_s": 0.0, "e2e_init_s": 0.0, "e2e_steps_s": 0.0,
               "e2e_read_s": 0.0, "attack_s": 0.0, "rows": 0, "utts_local": len(mine)}
        sampler = ClockSampler(local_rank)
        with torch.cuda.stream(stream):
            # warm the context (allocations, graph capture) outside every timed region
            fb = make_attacker(name, model, 2 * (K + W) + 64, 20261017)
            fb._nes_init(utts[mine[0]][:, None] if mine else utts[0][:, None], fb.max_iter)
            eng.nes_run(W)
            torch.cuda.synchronize()
            barrier()
            sampler.start()
            for u in mine:
                fb = make_attacker(name, model, 2 * (K + W) + 64, 20261017 + u)
                fb._nes_init(utts[u][:, None], fb.max_iter)
                eng.nes_run(W)
                torch.cuda.synchronize()
                launches0 = eng.kernel_launches()
                # (1) device-timed, L2 flushed before every timed iteration
                evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
                for a, b in evs:
                    flush_l2()
                    a.record(stream)
                    eng.nes_run(1)
                    b.record(stream)
                torch.cuda.synchronize()
                res["ms_flushed"] += sum(a.elapsed_time(b) for a, b in evs)
                res["launches"] += eng.kernel_launches() - launches0
                # (2) device-timed, back to back (the loop's real steady state: its working set is L2 resident)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.nes_run(K)
                e1.record(stream)
                torch.cuda.synchronize()
                res["ms_hot"] += e0.elapsed_time(e1)
                it_done, stopped = eng.nes_status()
                assert not stopped and it_done == W + 2 * K, (it_done, stopped)
                res["rows"] = eng.voiced_rows()
            barrier()
            sampler.stop_flag = True
            sampler.join(timeout=2.0)
            # (3) per-stage device times (events between kernels, no graph replay)
            eng.profile(True)
            eng.nes_run(min(K, 30))
            eng.nes_status()
            res["prof"] = eng.profile_read()
            eng.profile(False)
            if c["arch"] == "iv":
                st = eng.stats(0)
                res["active_frac"] = float((st["gamma"] != 0).mean())
            # (4) end to end through the public C-ABI with HOST buffers, host wall clock: one whole attack session per
            #     utterance = fb_nes_init (audio H2D from host memory) + K x [fb_nes_run(1) + fb_nes_status + log-row D2H]
            #     + fb_nes_read_adver (result D2H).  L2 is flushed (untimed) before every step.  Nothing is amortised or
            #     skipped; device buffers and the captured graph are reused from the warm context like in any second attack.
            if with_e2e:
                barrier()
                for u in mine:
                    fb2 = make_attacker(name, model, K, 777 + u)
                    audio_host = np.ascontiguousarray(utts[u][:, None])
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    eng2 = fb2._nes_init(audio_host, K)
                    t_init = time.perf_counter() - t0
                    t_steps = 0.0
                    for i in range(K):
                        flush_l2()
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        eng2.nes_run(1)
                        done, _ = eng2.nes_status()
                        eng2.nes_log(done)[-1]
                        t_steps += time.perf_counter() - t0
                    t0 = time.perf_counter()
                    eng2.nes_adver()
                    t_read = time.perf_counter() - t0
                    res["e2e_init_s"] += t_init
                    res["e2e_steps_s"] += t_steps
                    res["e2e_read_s"] += t_read
                    res["e2e_s"] += t_init + t_steps + t_read
                barrier()
                # (5) whole FakeBob.attack() calls (default iters_per_launch batching, no flush): informational
                t0 = time.perf_counter()
                for u in mine:
                    fb3 = make_attacker(name, model, K, 999 + u)
                    fb3.attack(utts[u], None, **attack_kwargs(name))
                res["attack_s"] = time.perf_counter() - t0
                barrier()
        res["clocks"] = sampler.summary()
        res["gmm_info"] = eng.gmm_info() if c["arch"] == "gmm" else None
        res["n_models"] = eng.n_models
        if c["arch"] == "gmm" and rank == 0:
            # deviation of the shipped contraction from the three-term one on the clean utterance (scores, live)
            a = model.score(utts[0])
            eng.set_delta_terms(3)
            b = model.score(utts[0])
            eng.set_delta_terms(int(os.environ.get("FAKEBOB_GMM_DELTA_TERMS", "0")))
            res["score_dev_vs_3term"] = float(np.abs(np.asarray(a) - np.asarray(b)).max())
        return res

    def line_for(name, res, K, W, peaks, peak_src):
        c = CONFIGS[name]
        ms_flushed = max_over_ranks(res["ms_flushed"])
        ms_hot = max_over_ranks(res["ms_hot"])
        e2e_s = max_over_ranks(res["e2e_s"])
        attack_s = max_over_ranks(res["attack_s"])
        launches = sum_over_ranks(res["launches"])
        if rank != 0:
            return None
        n_utts = c["utts"]
        # a step = one iteration of every utterance; ranks that shard utterances work concurrently
        ms_step = ms_flushed / K
        from fakebob_b200.sharding import pair_range
        p0, p1 = pair_range(c["S"] // 2, 0, world if c["shard"] == "pairs" else 1)
        roof, stage_ms = rooflines(name, res["prof"], res["rows"], res["n_models"], peaks, res.get("active_frac"), 1 + 2 * (p1 - p0))
        dom = max(roof.values(), key=lambda r: r["kernel_ms"]) if roof else None
        if dom:
            dom = dict(dom)
            dom["traffic"], dom["traffic_capture"] = load_traffic(dom["kernel"].split("<")[0])
            dom["peak_source"] = peak_src
        line = {
            "metric": "NES attack iters/sec", "value": K / (ms_flushed * 1e-3), "unit": "iters/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (fp16 hi/lo split operands, fp32 accumulate; f64 NES state)", "data": "synthetic",
            "config": config_dict(name),
            "timing": {"l2": "flushed (256 MiB memset, untimed) before every timed iteration" if not args.no_flush else "not flushed",
                       "rng": "philox (device)", "parallelism": ("%s sharded over %d rank(s)" % (c["shard"], world)) if multi else "single GPU",
                       "voiced_rows_per_iter_rank0": res["rows"], "utterance_iterations_per_step": n_utts,
                       "published_reference": "README.md:112-113: ~5 s / iteration (GMM-UBM), ~12 s (ivector-PLDA), hardware unspecified"},
            "value_hot_l2": K / (ms_hot * 1e-3),
            "gpu_launches": int(launches),
            "roofline": dom,
            "roofline_all": roof,
            "stage_ms": stage_ms,
            "clocks": res["clocks"],
        }
        if res["e2e_s"] > 0:
            h2d = 8 * N_SAMPLES * n_utts / K
            d2h = (16 + 8 * (4 + c["K"])) * n_utts + 8 * N_SAMPLES * n_utts / K
            line["e2e"] = {"value": K / e2e_s, "unit": "iters/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                           "init_ms": 1e3 * res["e2e_init_s"], "steps_ms": 1e3 * res["e2e_steps_s"], "read_ms": 1e3 * res["e2e_read_s"],
                           "how": "host wall clock of whole attack sessions through the C-ABI with host buffers: fb_nes_init (audio H2D) + "
                                  "%d x [fb_nes_run(1) + fb_nes_status + log row D2H] + fb_nes_read_adver (D2H); L2 flushed (untimed) before "
                                  "each step; warm context (buffers and graph reused, as for any attack after the first)" % K}
            line["e2e_attack_api"] = {"value": K / attack_s, "unit": "iters/s",
                                      "how": "FakeBob.attack(audio_host) call(s) of %d iterations, default iters_per_launch, host wall clock" % K}
        if res.get("gmm_info"):
            line["gmm_contraction"] = dict(res["gmm_info"], score_deviation_vs_three_term=res.get("score_dev_vs_3term"),
                                           note="speaker slots scored as slot 0 + x.(w_m - w_0) with delta_terms fp16 products; deviation "
                                                "from the oracle at this size is asserted in tests/test_gpu_fullsize.py")
        return line

    name = args.config
    extra = (name == "C2" and world == 1 and not args.no_extra and os.environ.get("FB_BENCH_SKIP_IV") is None)
    tree = build_trees([name] + (["C3"] if extra else []))
    peaks, peak_src = load_peaks()
    K, W = args.steps, args.warmup
    res = measure(name, tree, K, W)
    line = line_for(name, res, K, W, peaks, peak_src)
    if extra:
        try:
            r3 = measure("C3", tree, min(K, 50), min(W, 10))
            l3 = line_for("C3", r3, min(K, 50), min(W, 10), peaks, peak_src)
            if rank == 0:
                line["extra_config_C3"] = {k: l3[k] for k in ("value", "ms_per_step", "value_hot_l2", "e2e", "roofline", "roofline_all",
                                                             "stage_ms", "config", "gpu_launches") if k in l3}
        except Exception as e:                           # informational only: never break the contract line
            if rank == 0:
                line["extra_config_C3"] = {"error": repr(e)[:300]}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            t0 = time.perf_counter()
            cpu_tree = tree
            ips, sample, cores, loop = time_cpu(name, cpu_tree, tree["root"], steps=2, warmup=1, budget_s=25.0)
            line["cpu_baseline"] = {"value": ips, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample, "loop": loop,
                                    "wall_s": time.perf_counter() - t0}
        print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
