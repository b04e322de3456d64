#!/usr/bin/env python
"""Benchmark of the FAKEBOB NES attack hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], "C2"): gmm_OSI untargeted NES attack, UBM + 5 MAP-adapted speaker
GMMs with 2048 mixtures, samples_per_draw = 50, one 5 s @ 16 kHz utterance, epsilon = 0.002, early stop
disabled (unreachable threshold) so every iteration does the full work.  One "step" = one NES iteration
(FAKEBOB.py:168-214): 51 audios perturbed, quantised, MFCC'd, VAD'd, delta/CMN'd, scored against 6 GMMs,
loss, gradient estimate, momentum/sign/clip update.  Data and models are synthetic (fakebob_b200/synth.py).

Prints ONE JSON line (rank 0).  See the module docstring of DESIGN.md section "Measurement" for the keys.
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

S_DRAW = 50
N_SAMPLES = 80000
N_SPEAKERS = 5
N_MIX = 2048
EPSILON = 0.002
# dram__bytes_read.sum + dram__bytes_write.sum of gmm_umma_kernel at this workload, one `ncu --set full` capture
# (profiles/r01_ncu_iter_v5_summary.txt): 19.53 MB read (the fp16 operand image written by feats_kernel; the 4.5 MB W image
# stays in L2), ~0 written (partials stay in L2).  The kernel is tensor-pipe bound; traffic is reported for completeness.
GMM_DRAM_BYTES_PER_LAUNCH = 19533056
WORKLOAD = "C2: gmm_OSI untargeted, UBM+5 spk x 2048 mix, samples_per_draw=50, 5 s @ 16 kHz, eps=0.002"


def env_int(name, default):
    return int(os.environ.get(name, default))


# --------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock / power / throttle reasons during the timed region (B200_PROFILING.md recipe), sampled through NVML
    every 50 ms (falls back to the nvidia-smi query line when pynvml is unavailable).  The sampler is stopped before the
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


# --------------------------------------------------------------------------------------------------
def build_workload_gpu(root, device):
    """Synthetic pre-models/ + model/ tree whose features come from the CUDA front-end."""
    from fakebob_b200 import synth
    from fakebob_b200.engine import GmmEngine
    r = np.random.default_rng(0)
    dummy = {"weights": np.full(128, 1 / 128, np.float32), "means_invvars": r.standard_normal((128, 72)).astype(np.float32),
             "inv_vars": np.ones((128, 72), np.float32), "gconsts": np.zeros(128, np.float32)}
    fe = GmmEngine([dummy], device=device)
    tree = synth.build_gmm_tree(root, fe.features, n_speakers=N_SPEAKERS, C=N_MIX, n_ubm_utts=64, n_samples=N_SAMPLES)
    fe.close()
    return tree


def build_workload_cpu(root):
    from fakebob_b200 import synth
    from oracle import kaldi_feats as kf
    return synth.build_gmm_tree(root, kf.voiced_features, n_speakers=N_SPEAKERS, C=N_MIX, n_ubm_utts=64, n_samples=N_SAMPLES)


def oracle_model(tree):
    from fakebob_b200 import kaldi_io
    from oracle.diag_gmm import DiagGmm
    from oracle.scorers import OracleGmmOSI

    def lg(p):
        g = kaldi_io.read_diag_gmm(p)
        return DiagGmm(g["weights"], g["means_invvars"], g["inv_vars"], g["gconsts"])
    return OracleGmmOSI(lg(tree["ubm"]), [lg(m[2]) for m in tree["models"]])


def time_oracle(tree, audio, steps, warmup, budget_s=200.0):
    """Times the CPU restatement (oracle NES loop + oracle Kaldi scoring) on the host cores.
    Returns (iters_per_sec, sample description, cores)."""
    from oracle.nes import OracleFakeBob, PhiloxNoise
    model = oracle_model(tree)
    cores = os.cpu_count() or 1
    try:
        import torch
        torch.set_num_threads(cores)
    except Exception:
        pass
    s_draw = S_DRAW
    fb = OracleFakeBob("OSI", "untargeted", model, max_iter=1, samples_per_draw=s_draw, noise_fn=PhiloxNoise(1))
    t0 = time.perf_counter()
    fb.attack(audio, None, threshold=1e3)
    t_first = time.perf_counter() - t0
    scale = 1.0
    note = "full NES iteration (S=50, 51 audios x 6 GMMs x 2048 mix)"
    if (steps + warmup) * t_first > budget_s:
        s_draw = 10
        scale = (s_draw + 1) / float(S_DRAW + 1)
        note = "NES iteration with S=10 (11 audios), rate scaled by 11/51 to S=50"
    n_warm = max(warmup - (1 if s_draw == S_DRAW else 0), 0)
    if n_warm:
        OracleFakeBob("OSI", "untargeted", model, max_iter=n_warm, samples_per_draw=s_draw,
                      noise_fn=PhiloxNoise(2)).attack(audio, None, threshold=1e3)
    fb = OracleFakeBob("OSI", "untargeted", model, max_iter=steps, samples_per_draw=s_draw, noise_fn=PhiloxNoise(1))
    t0 = time.perf_counter()
    fb.attack(audio, None, threshold=1e3)
    per_iter = (time.perf_counter() - t0) / steps
    return scale / per_iter, "%d x %s" % (steps, note), cores


def run_reference(args, rank):
    if rank != 0:
        return
    from fakebob_b200 import synth
    root = tempfile.mkdtemp(prefix="fakebob_ref_")
    tree = build_workload_cpu(root)
    audio = synth.synth_utterance(0, 0, N_SAMPLES)
    t0 = time.perf_counter()
    ips, sample, cores = time_oracle(tree, audio, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "NES attack iters/sec", "value": ips, "unit": "iters/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / ips, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement of the reference path (oracle/: numpy Kaldi arithmetic driven by "
                   "the restated FAKEBOB.py loop); the real Kaldi-subprocess path cannot run offline (no Kaldi, no pre-models). "
                   "Omits the reference's fork/exec and wav/ark/text I/O, so it is faster than the real reference."},
        "cpu_baseline": {"value": ips, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def bench_ivector(root, ubm_params, pre_model_dir, device, audio, iters=100):
    """Informational: BASELINE.json configs[2] (iv_SV, 2048-mix full UBM, 400-dim i-vector, LDA 200 + PLDA, S=50)."""
    from fakebob_b200 import synth
    from fakebob_b200.engine import IvectorEngine
    from fakebob_b200.FAKEBOB import FakeBob
    from fakebob_b200.ivector_PLDA_SV import iv_SV
    synth.build_ivector_params(root, ubm_params, R=400, L=200)
    eng = IvectorEngine(pre_model_dir, device=device)
    spk = synth.build_ivector_speakers(root, lambda w: eng.extract_ivectors([np.ascontiguousarray(w, dtype=np.int16)])[0],
                                       lambda enrolled, test: np.arange(len(test), dtype=np.float64)[:, None] + np.zeros((1, len(enrolled))),
                                       n_speakers=1, n_samples=N_SAMPLES, n_znorm_utts=2)
    eng.close()
    model = iv_SV(os.path.join(root, "iv-sv"), spk["models"][0], pre_model_dir=pre_model_dir, device=device)
    fb = FakeBob("SV", "untargeted", model, max_iter=30, samples_per_draw=S_DRAW, seed=1, verbose=False)
    fb.attack(audio, None, threshold=1e9)
    fb = FakeBob("SV", "untargeted", model, max_iter=iters, samples_per_draw=S_DRAW, seed=1, verbose=False)
    t0 = time.perf_counter()
    fb.attack(audio, None, threshold=1e9)
    dt = time.perf_counter() - t0
    e = model._engine
    e.profile(True)
    fb = FakeBob("SV", "untargeted", model, max_iter=10, samples_per_draw=S_DRAW, seed=1, verbose=False)
    fb.attack(audio, None, threshold=1e9)
    prof = e.profile_read()
    e.profile(False)
    return {"workload": "C3: iv_SV, 2048-mix full UBM, 400-dim i-vector, LDA 200 + PLDA, samples_per_draw=50, 5 s @ 16 kHz",
            "iters_per_s_attack_api": fb.max_iter and iters / dt, "ms_per_iter": dt / iters * 1e3,
            "published_reference": "README.md:112-113: ~12 s / iteration (ivector-PLDA)",
            "stage_ms": {k: v[0] / max(v[1], 1) for k, v in prof.items() if v[1]}}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    from fakebob_b200 import synth
    from fakebob_b200.gmm_ubm_OSI import gmm_OSI
    from fakebob_b200.FAKEBOB import FakeBob

    # ---- workload: rank 0 builds the synthetic tree, the others load it
    holder = [None]
    if rank == 0:
        root = tempfile.mkdtemp(prefix="fakebob_bench_")
        tree = build_workload_gpu(root, local_rank)
        holder[0] = {k: tree[k] for k in ("pre_model_dir", "model_dir", "ubm", "spk_ids", "models")}
        ubm_params, bench_root = tree["ubm_params"], root
    if multi:
        dist.broadcast_object_list(holder, src=0)
    tree = holder[0]
    model = gmm_OSI(os.path.join(tempfile.mkdtemp(prefix="fakebob_grp_"), "gmm-OSI-untargeted"), tree["models"], tree["ubm"],
                    pre_model_dir=tree["pre_model_dir"], threshold=0.0, device=local_rank)
    eng = model._engine
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    if multi:
        eng.comm_init_from_torch()
    audio = synth.synth_utterance(0, 0, N_SAMPLES)
    K, W = args.steps, args.warmup
    theta = 1e3                                           # unreachable: early stop never fires

    fb = FakeBob("OSI", "untargeted", model, epsilon=EPSILON, max_iter=2 * (K + W) + 64, samples_per_draw=S_DRAW,
                 seed=20261017, verbose=False)
    fb.threshold = theta
    eng = fb._nes_init(audio[:, None], fb.max_iter)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.zero_()

    sampler = ClockSampler(local_rank)
    with torch.cuda.stream(stream):
        eng.nes_run(W)
        torch.cuda.synchronize()
        launches0 = eng.kernel_launches()
        barrier()
        torch.cuda.synchronize()
        sampler.start()
        # ---- (1) device-timed, L2 flushed before every timed iteration
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for a, b in evs:
            if not args.no_flush:
                flush_l2()
            a.record(stream)
            eng.nes_run(1)
            b.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms_flushed = sum(a.elapsed_time(b) for a, b in evs)
        launches = eng.kernel_launches() - launches0
        # ---- (2) device-timed, back-to-back (the loop's real steady state: working set is L2 resident)
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        eng.nes_run(K)
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        ms_hot = e0.elapsed_time(e1)
        sampler.stop_flag = True
        sampler.join(timeout=2.0)
        it_done, stopped = eng.nes_status()
        assert not stopped and it_done == W + 2 * K, (it_done, stopped)
        rows = eng.voiced_rows()
        # ---- (3) per-stage device times (events between kernels, no graph replay)
        eng.profile(True)
        n_prof = min(K, 50)
        eng.nes_run(n_prof)
        prof = eng.profile_read()
        eng.profile(False)

        # ---- (4) end to end through the public C-ABI step with host buffers: per step, L2 flush (untimed), then
        #          fb_nes_run(1) + fb_nes_status + log-row read; audio upload (fb_nes_init) timed once and amortised
        K_e2e = K
        t_init0 = time.perf_counter()
        fb2 = FakeBob("OSI", "untargeted", model, epsilon=EPSILON, max_iter=K_e2e + W, samples_per_draw=S_DRAW,
                      seed=20261017, verbose=False)
        fb2.threshold = theta
        eng2 = fb2._nes_init(audio[:, None], fb2.max_iter)
        torch.cuda.synchronize()
        t_init = time.perf_counter() - t_init0
        eng2.nes_run(W)
        eng2.nes_status()
        barrier()
        t_steps = 0.0
        for i in range(K_e2e):
            if not args.no_flush:
                flush_l2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng2.nes_run(1)
            done, _ = eng2.nes_status()
            eng2.nes_log(done)[-1]
            t_steps += time.perf_counter() - t0
        barrier()
        e2e_s = t_steps + t_init
        # ---- (5) whole FakeBob.attack() call (default batching, no flush): informational
        fb3 = FakeBob("OSI", "untargeted", model, epsilon=EPSILON, max_iter=K, samples_per_draw=S_DRAW, seed=1, verbose=False)
        barrier()
        t0 = time.perf_counter()
        fb3.attack(audio, None, threshold=theta)
        t_attack = time.perf_counter() - t0
        barrier()

    def max_over_ranks(x):
        if not multi:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_flushed = max_over_ranks(ms_flushed)
    ms_hot = max_over_ranks(ms_hot)
    e2e_s = max_over_ranks(e2e_s)
    t_attack = max_over_ranks(t_attack)
    clocks = sampler.summary()

    if rank == 0:
        peaks = {}
        for p in (os.path.join(ROOT, "MEASURED_PEAKS.json"),):
            if os.path.exists(p):
                with open(p) as f:
                    peaks = json.load(f)
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)" if peaks else \
            "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        n_models = N_SPEAKERS + 1
        gmm_ms, gmm_n = prof["gmm"]
        gmm_ms_avg = gmm_ms / max(gmm_n, 1)
        flops = 2.0 * rows * N_MIX * 144 * n_models          # per launch, this rank's rows
        achieved = flops / (gmm_ms_avg * 1e-3) / 1e12 if gmm_ms_avg > 0 else 0.0
        value = K / (ms_flushed * 1e-3)
        line = {
            "metric": "NES attack iters/sec", "value": value, "unit": "iters/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_flushed / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (fp16 hi/lo split operands, fp32 accumulate; f64 NES state)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rng": "philox (device)", "parallelism": "antithetic pairs sharded over %d rank(s), "
                       "1 ncclAllReduce(f64, N+S+1+K) per iteration" % world if multi else "single GPU",
                       "l2": "flushed (256 MiB memset, untimed) before every timed iteration" if not args.no_flush else "not flushed",
                       "voiced_rows_per_iter": rows, "published_reference": "README.md:112-113: ~5 s / iteration (GMM-UBM), hardware unspecified"},
            "value_hot_l2": K / (ms_hot * 1e-3),
            "e2e": {"value": K_e2e / e2e_s, "unit": "iters/s", "h2d_bytes_per_step": int(8 * N_SAMPLES / K_e2e),
                    "d2h_bytes_per_step": 16 + 8 * (4 + N_SPEAKERS) * 1,
                    "how": "per step: fb_nes_run(1) + fb_nes_status + log read through the C-ABI, host wall clock, L2 flushed "
                           "before each step; audio upload (fb_nes_init) included once"},
            "e2e_attack_api": {"value": K / t_attack, "unit": "iters/s",
                               "how": "one FakeBob.attack(audio_host) call of %d iterations, default iters_per_launch, host wall clock" % K},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "gmm_umma_kernel", "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved / peak_tf if peak_tf else None, "traffic": GMM_DRAM_BYTES_PER_LAUNCH,
                         "traffic_unit": "bytes/launch (dram read + write, profiles/r01_ncu_iter_v5_summary.txt)", "peak_source": peak_src,
                         "algorithmic_flops_per_launch": flops, "kernel_ms": gmm_ms_avg,
                         "executed_flops_per_launch": 3.0 * flops,
                         "note": "algorithmic = 2*rows*C*(2D) per model (Kaldi's two sgemv per frame); the kernel executes 3x that "
                                 "in fp16 MMAs (hi.hi + lo.hi + hi.lo) to keep fp32-class accuracy"},
            "stage_ms": {k: (v[0] / max(v[1], 1)) for k, v in prof.items()},
            "clocks": clocks,
        }
        if world == 1 and os.environ.get("FB_BENCH_SKIP_IV") is None:
            try:
                line["extra_config_C3"] = bench_ivector(bench_root, ubm_params, tree["pre_model_dir"], local_rank, audio)
            except Exception as e:                       # informational only: never break the contract line
                line["extra_config_C3"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            t0 = time.perf_counter()
            ips, sample, cores = time_oracle(tree, audio, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": ips, "unit": "iters/s", "cores": cores, "kind": "port", "sample": sample,
                                    "wall_s": time.perf_counter() - t0}
        print(json.dumps(line), flush=True)
    if multi:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
