"""Host-side logic that needs no GPU: Kaldi file formats, config parsing, input conventions, the C-ABI library
loading and exporting every declared symbol, loud failure without a GPU, and the S-sharding scheme over gloo."""
import ctypes
import os
import pickle
import re
import sys

import numpy as np
import pytest

from fakebob_b200 import kaldi_io, synth
from fakebob_b200.config import FeatureConfig, load_feature_config
from fakebob_b200.engine import to_audio_list

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_kaldi_io_round_trips(tmp_path):
    r = np.random.default_rng(0)
    C, D, R, L = 6, 72, 10, 5
    w = r.dirichlet(np.ones(C)).astype(np.float32)
    miv = r.standard_normal((C, D)).astype(np.float32)
    iv = r.uniform(0.5, 2, (C, D)).astype(np.float32)
    p = str(tmp_path / "a.gmm")
    kaldi_io.write_diag_gmm(p, w, miv, iv)
    g = kaldi_io.read_diag_gmm(p)
    assert np.array_equal(g["weights"], w) and np.array_equal(g["means_invvars"], miv) and np.array_equal(g["inv_vars"], iv)
    assert np.allclose(g["gconsts"], kaldi_io.diag_gconsts(w, miv, iv))
    with open(p, "rb") as f:
        assert f.read(12) == b"\0B<DiagGMM> "
    ic = np.stack([np.eye(D) * (1 + i) + 0.01 for i in range(C)])
    p = str(tmp_path / "f.ubm")
    kaldi_io.write_full_gmm(p, w, miv, ic, np.zeros(C))
    fg = kaldi_io.read_full_gmm(p)
    assert fg["inv_covars"].shape == (C, D, D) and np.allclose(fg["inv_covars"], ic)
    M = r.standard_normal((C, D, R))
    p = str(tmp_path / "final.ie")
    kaldi_io.write_ivector_extractor(p, w, M, ic, 100.0)
    ie = kaldi_io.read_ivector_extractor(p)
    assert np.array_equal(ie["M"], M) and ie["prior_offset"] == 100.0 and ie["w"].size == 0 and np.allclose(ie["sigma_inv"], ic)
    p = str(tmp_path / "plda")
    kaldi_io.write_plda(p, np.arange(L), np.eye(L), np.arange(L) + 1.0)
    pl = kaldi_io.read_plda(p)
    assert np.array_equal(pl["psi"], np.arange(L) + 1.0) and pl["transform"].shape == (L, L)
    for binary in (True, False):
        p = str(tmp_path / ("m%d.mat" % binary))
        m = r.standard_normal((3, 4)).astype(np.float32)
        kaldi_io.write_matrix(p, m, binary=binary)
        assert np.allclose(kaldi_io.read_matrix(p), m, rtol=1e-6)
        p = str(tmp_path / ("v%d.vec" % binary))
        kaldi_io.write_vector(p, m[0], binary=binary)
        assert np.allclose(kaldi_io.read_vector(p), m[0], rtol=1e-6)


def test_text_ark_scp_targets(tmp_path):
    v = {"a-1": np.array([1.5, -2.25, 3.0]), "b-2": np.array([0.1234567891, 7.0, -8.0])}
    t = kaldi_io.write_text_vector_ark(str(tmp_path / "ivector.1.ark"), list(v.items()))
    assert re.match(r".*ivector\.1\.ark:\d+$", t["b-2"])
    assert np.allclose(kaldi_io.read_vector(t["a-1"]), v["a-1"])
    assert np.allclose(kaldi_io.read_vector(t["b-2"]), [0.1234568, 7.0, -8.0])       # 7 significant digits ('ark,t')


def test_feature_config_from_pre_models(tmp_path):
    synth.write_conf(str(tmp_path))
    cfg = load_feature_config(str(tmp_path))
    assert (cfg.num_mel_bins, cfg.num_ceps, cfg.snip_edges, cfg.high_freq) == (30, 24, False, 7600.0)
    assert (cfg.delta_window, cfg.delta_order, cfg.vad_energy_threshold) == (3, 2, 5.5)
    assert cfg.num_frames(80000) == 500 and cfg.feat_dim == 72
    cfg.check_supported()
    bad = FeatureConfig(num_ceps=13)
    with pytest.raises(ValueError):
        bad.check_supported()
    assert load_feature_config(str(tmp_path / "missing")).num_mel_bins == 30


def test_feature_config_rejects_options_the_kernels_do_not_implement(tmp_path):
    """mfcc.conf keys that change the features must not be dropped silently (the reference hands the file to Kaldi,
    gmm_ubm_kaldiHelper.py:138): every Kaldi MFCC option is parsed, unsupported values raise, unknown keys warn."""
    import warnings
    for line, ok in (("--use-energy=false", False), ("--raw-energy=false", False), ("--energy-floor=1.0", False),
                     ("--window-type=hamming", False), ("--remove-dc-offset=false", False), ("--dither=1.0", False),
                     ("--htk-compat=true", False), ("--use-energy=true", True), ("--dither=0", True), ("--window-type=povey", True)):
        d = tmp_path / line.strip("-").replace("=", "_")
        synth.write_conf(str(d))
        with open(str(d / "conf" / "mfcc.conf"), "a") as f:
            f.write(line + "\n")
        cfg = load_feature_config(str(d))
        if ok:
            cfg.check_supported()
        else:
            with pytest.raises(ValueError):
                cfg.check_supported()
    d = tmp_path / "unknown"
    synth.write_conf(str(d))
    with open(str(d / "conf" / "mfcc.conf"), "a") as f:
        f.write("--vtln-low=100\n")
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        load_feature_config(str(d)).check_supported()
    assert any("vtln-low" in str(x.message) for x in w)


def test_audio_input_conventions():
    a = np.linspace(-0.5, 0.5, 100)
    for arr in (a, a[:, None], a[None, :]):
        lst = to_audio_list(arr)
        assert len(lst) == 1 and lst[0].shape == (100,) and lst[0].dtype == np.int16
    lst = to_audio_list(np.stack([a, -a], axis=1))
    assert len(lst) == 2 and np.array_equal(lst[1], (-a * 32768).astype(np.int16))
    lst = to_audio_list([a[:10], np.arange(7, dtype=np.int16)])
    assert [x.shape[0] for x in lst] == [10, 7] and np.array_equal(lst[1], np.arange(7))
    assert to_audio_list(np.array([1.9 / 32768, -1.9 / 32768]))[0].tolist() == [1, -1]


def test_library_exports_every_declared_symbol():
    from fakebob_b200 import _lib
    lib = _lib.load()
    with open(os.path.join(ROOT, "include", "fakebob_b200.h")) as f:
        declared = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", f.read()))
    assert declared, "no declarations found"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), name
        assert name in _lib.EXPORTS, "no ctypes prototype for " + name
    assert lib.fb_version() >= 100


def test_fails_loudly_without_gpu_or_device_model():
    import torch
    from fakebob_b200 import _lib
    from fakebob_b200.FAKEBOB import FakeBob
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        rc = _lib.load().fb_ctx_create(0, ctypes.byref(h))
        assert rc < 0 and b"no CPU fallback" in _lib.load().fb_last_error()
        from fakebob_b200.engine import GmmEngine
        with pytest.raises(_lib.FakebobLibraryError):
            GmmEngine([{"weights": np.ones(128, np.float32) / 128, "means_invvars": np.zeros((128, 72), np.float32),
                        "inv_vars": np.ones((128, 72), np.float32), "gconsts": np.zeros(128, np.float32)}])

    class Stub:
        def score(self, a):
            return 0.0
    with pytest.raises(TypeError):
        FakeBob("SV", "untargeted", Stub())                     # not a scorer: no make_decisions()

    class BlackBox(Stub):
        def make_decisions(self, a):
            return -1, 0.0
    if not torch.cuda.is_available():
        # a black-box model is accepted, but its attack still needs the device (NES state, noise, update): no CPU fallback
        fb = FakeBob("SV", "untargeted", BlackBox(), max_iter=1, samples_per_draw=2, verbose=False)
        with pytest.raises(_lib.FakebobLibraryError):
            fb.attack(np.zeros(1600), None, threshold=1.0)


def test_synthetic_tree_layout(small_tree):
    t = small_tree
    assert os.path.exists(os.path.join(t["pre_model_dir"], "final.dubm"))
    assert os.path.exists(os.path.join(t["pre_model_dir"], "conf", "mfcc.conf"))
    with open(os.path.join(t["model_dir"], t["spk_ids"][0] + ".gmm"), "rb") as f:
        m = pickle.load(f)
    assert len(m) == 5 and m[0] == t["spk_ids"][0] and os.path.isabs(m[2]) and m[2].endswith("-identity.gmm")
    g = kaldi_io.read_diag_gmm(m[2])
    u = kaldi_io.read_diag_gmm(t["ubm"])
    assert np.array_equal(g["inv_vars"], u["inv_vars"]) and not np.array_equal(g["means_invvars"], u["means_invvars"])


# ---- multi-GPU scheme on CPU: pairs sharded over 2 gloo ranks, one all-reduce of [grad | losses | scores] -------
def _shard_worker(rank, world, port, q):
    import sys
    import torch.distributed as dist
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from stub_scorer import StubScorer, make_audio
    from fakebob_b200.sharding import pair_range, global_column
    from oracle.nes import margin_loss
    from oracle import philox
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, S, K, sigma, seed, it = 4000, 10, 3, 1e-3, 5, 2
    model = StubScorer(K, N)
    audio = make_audio(1)[:, None]
    p0, p1 = pair_range(S // 2, rank, world)
    noise = philox.normal_noise(seed, it, N, S // 2)[:, p0:p1]
    cols = [0] if rank == 0 else []
    batch = [audio] if rank == 0 else []
    batch += [sigma * noise + audio, sigma * (-1.0 * noise) + audio]
    cols += [global_column(S // 2, p0 + j, +1) for j in range(p1 - p0)] + [global_column(S // 2, p0 + j, -1) for j in range(p1 - p0)]
    scores = model.score(np.concatenate(batch, axis=1))
    loss = margin_loss(scores, "OSI", "untargeted", 2.5, 0.0).reshape(-1)
    red = np.zeros(N + S + 1 + K)
    red[N + np.array(cols)] = loss
    if rank == 0:
        red[N + S + 1:] = scores[0]
        loc = loss[1:]
    else:
        loc = loss
    npairs = p1 - p0
    red[:N] = (loc[:npairs] * noise).sum(axis=1) + (loc[npairs:] * (-1.0 * noise)).sum(axis=1)
    t = torch.from_numpy(red)
    dist.all_reduce(t)
    if rank == 0:
        q.put(t.numpy().copy())
    dist.destroy_process_group()


def test_sharded_gradient_allreduce_matches_single_process_gloo():
    import sys
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from stub_scorer import StubScorer, make_audio
    from oracle.nes import OracleFakeBob, PhiloxNoise
    from fakebob_b200.sharding import pair_range
    assert [pair_range(25, r, 8) for r in range(8)] == [(0, 3), (3, 6), (6, 9), (9, 12), (12, 15), (15, 18), (18, 21), (21, 25)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + (os.getpid() % 200)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    red = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    N, S, K = 4000, 10, 3
    fb = OracleFakeBob("OSI", "untargeted", StubScorer(K, N), samples_per_draw=S, sigma=1e-3, noise_fn=PhiloxNoise(5))
    fb.threshold = 2.5
    fb.draws = 2
    final_loss, grad, adver_loss, score = fb.get_grad(make_audio(1))
    assert np.allclose(red[:N] / S / 1e-3, grad[:, 0], rtol=1e-12, atol=1e-12)
    assert red[N] == adver_loss[0] and np.isclose(red[N + 1:N + 1 + S].mean(), final_loss, rtol=1e-14)
    assert np.array_equal(red[N + S + 1:], score)


def test_set_threshold_matches_reference_search():
    """fakebob_b200.evaluate.set_threshold against a direct restatement of the loop at test.py:46-71."""
    from fakebob_b200.evaluate import set_threshold

    def slow(score_target, score_untarget):
        best, best_d, best_far, best_frr = 0.0, np.inf, 0.0, 0.0
        for thr in score_target:
            frr = np.count_nonzero(score_target < thr) * 100 / score_target.size
            far = np.count_nonzero(score_untarget >= thr) * 100 / score_untarget.size
            if abs(frr - far) < best_d:
                best, best_d, best_far, best_frr = thr, abs(frr - far), far, frr
        return best, best_frr, best_far

    r = np.random.default_rng(5)
    for n_t, n_u in ((1, 1), (7, 3), (50, 200), (200, 50)):
        st = r.normal(1.0, 1.0, n_t)
        su = r.normal(-1.0, 1.0, n_u)
        st[::5] = np.round(st[::5], 1)                      # ties
        su[::4] = np.round(su[::4], 1)
        assert set_threshold(st, su) == pytest.approx(slow(st, su), abs=0)
    assert set_threshold([0.5, 0.5, 2.0], [0.5, -1.0])[0] == 0.5


def test_evaluation_helpers_with_stub_scorers():
    """csi_accuracy / sv_error_rates / osi_error_rates (test.py sections) on scorers with known scores."""
    from fakebob_b200 import evaluate as ev

    class StubCsi:
        spk_ids = ["a", "b", "c"]

        def make_decisions(self, audios, **kw):
            return [int(np.argmax(a[:3])) for a in audios], None

    audios = [np.array([0.1, 0.9, 0.0]), np.array([0.7, 0.2, 0.1]), np.array([0.1, 0.2, 0.9]), np.array([0.5, 0.4, 0.1])]
    assert ev.csi_accuracy(StubCsi(), audios, [1, 0, 2, 1]) == 75.0

    class StubSv:
        threshold = 0.0

        def score(self, audios, **kw):
            return np.array([float(a[0]) for a in audios])

    sv = StubSv()
    tgt = [np.array([s]) for s in (2.0, 1.5, 0.4, 3.0)]
    imp = [np.array([s]) for s in (0.5, -1.0, 0.1, 1.6)]
    r = ev.sv_error_rates(sv, tgt, imp)
    assert r["threshold"] == sv.threshold and r["frr"] == 25.0 and r["far"] == 25.0
    r2 = ev.sv_error_rates(sv, tgt, imp, threshold=10.0)
    assert r2 == {"threshold": 10.0, "frr": 100.0, "far": 0.0}

    class StubOsi:
        threshold = 0.0

        def score(self, audios, **kw):
            return np.stack([a[:2] for a in audios])

    osi = StubOsi()
    enrolled = [np.array([2.0, 0.0]), np.array([0.1, 3.0]), np.array([1.0, 1.2]), np.array([-1.0, -2.0])]
    illegal = [np.array([-3.0, -2.5]), np.array([0.5, 2.5])]
    r = ev.osi_error_rates(osi, enrolled, [0, 1, 0, 1], illegal, threshold=0.0)
    assert r == {"threshold": 0.0, "frr": 25.0, "ier": 25.0, "far": 50.0}


def test_enrolment_directory_conventions(tmp_path):
    """utt_id = file name up to the first '.', spk_id = utt_id up to the first '-' (build_spk_models.py:81-83); 16 kHz 16-bit wav only."""
    from scipy.io import wavfile
    from fakebob_b200 import build_spk_models as bsm
    d = tmp_path / "enrollment-set"
    d.mkdir()
    for name, fs in (("1580-141083-0000.wav", 16000), ("61-70968-0001.wav", 16000)):
        wavfile.write(str(d / name), fs, (np.arange(1600) % 100).astype(np.int16))
    utt, spk, paths = bsm.list_audio_dir(str(d))
    assert utt == ["1580-141083-0000", "61-70968-0001"] and spk == ["1580", "61"]
    a = bsm.read_wav_int16(paths[0])
    assert a.dtype == np.int16 and a.shape == (1600,)
    wavfile.write(str(d / "bad-1.wav"), 8000, np.zeros(800, np.int16))
    with pytest.raises(ValueError):
        bsm.read_wav_int16(str(d / "bad-1.wav"))
    wavfile.write(str(d / "bad-2.wav"), 16000, np.zeros(800, np.float32))
    with pytest.raises(ValueError):
        bsm.read_wav_int16(str(d / "bad-2.wav"))


def test_utterance_sharding_covers_every_utterance_once():
    from fakebob_b200.sharding import attack_many, utterance_range

    for n, world in ((32, 8), (5, 2), (3, 8), (0, 4)):
        got = []
        for r in range(world):
            lo, hi = utterance_range(n, r, world)
            assert 0 <= lo <= hi <= n
            got += list(range(lo, hi))
        assert got == list(range(n))

    class StubAttacker:
        seed, draws = 7, 0

        def attack(self, audio, checkpoint_path, **kw):
            return (self.seed, self.draws, float(audio.sum()), kw), 1

    audios = [np.full(4, i, dtype=np.float64) for i in range(5)]
    a = attack_many(StubAttacker, audios, rank=1, world=2, threshold=lambda u: 10.0 + u, target=2)
    assert sorted(a) == [2, 3, 4]
    whole = attack_many(StubAttacker, audios, threshold=lambda u: 10.0 + u, target=2)
    for u in a:
        assert a[u] == whole[u]                                   # same per-utterance seed whatever the sharding
        assert a[u][0][3] == {"threshold": 10.0 + u, "target": 2}
    assert len({whole[u][0][0] for u in whole}) == 5              # distinct streams per utterance


def test_dropin_modules_resolve_reference_imports():
    """With fakebob_b200/dropin first on sys.path, the import lines of the reference's attackMain.py (:15-21) and test.py
    (:8-13) resolve to this package's classes, whose constructors and methods take the reference's arguments."""
    import inspect
    import subprocess
    code = (
        "from FAKEBOB import FakeBob\n"
        "from gmm_ubm_CSI import gmm_CSI\nfrom gmm_ubm_OSI import gmm_OSI\nfrom gmm_ubm_SV import gmm_SV\n"
        "from ivector_PLDA_CSI import iv_CSI\nfrom ivector_PLDA_OSI import iv_OSI\nfrom ivector_PLDA_SV import iv_SV\n"
        "import fakebob_b200.FAKEBOB as F, fakebob_b200.gmm_scorers as G, fakebob_b200.iv_scorers as I\n"
        "assert FakeBob is F.FakeBob and gmm_CSI is G.gmm_CSI and gmm_OSI is G.gmm_OSI and gmm_SV is G.gmm_SV\n"
        "assert iv_CSI is I.iv_CSI and iv_OSI is I.iv_OSI and iv_SV is I.iv_SV\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "fakebob_b200", "dropin"), ROOT]))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr
    from fakebob_b200.FAKEBOB import FakeBob
    from fakebob_b200.gmm_scorers import gmm_CSI, gmm_OSI, gmm_SV
    from fakebob_b200.iv_scorers import iv_CSI, iv_OSI, iv_SV

    def names(fn):
        return [p for p in inspect.signature(fn).parameters if p != "self"]
    # reference signatures; keyword-only extras may follow
    assert names(FakeBob.__init__)[:13] == ["task", "attack_type", "model", "adver_thresh", "epsilon", "max_iter", "max_lr", "min_lr",
                                            "samples_per_draw", "sigma", "momentum", "plateau_length", "plateau_drop"]
    assert names(FakeBob.attack) == ["audio", "checkpoint_path", "threshold", "true", "target", "fs", "bits_per_sample", "n_jobs", "debug"]
    assert names(FakeBob.estimate_threshold) == ["audio", "fs", "bits_per_sample", "n_jobs", "debug"]
    assert names(gmm_CSI.__init__)[:3] == ["group_id", "model_list", "pre_model_dir"]
    assert names(gmm_OSI.__init__)[:5] == ["group_id", "model_list", "ubm", "pre_model_dir", "threshold"]
    assert names(gmm_SV.__init__)[:5] == ["spk_id", "model", "ubm", "pre_model_dir", "threshold"]
    assert names(iv_CSI.__init__)[:3] == ["group_id", "model_list", "pre_model_dir"]
    assert names(iv_OSI.__init__)[:4] == ["group_id", "model_list", "pre_model_dir", "threshold"]
    assert names(iv_SV.__init__)[:4] == ["spk_id", "model", "pre_model_dir", "threshold"]
    for cls in (gmm_CSI, gmm_OSI, gmm_SV, iv_CSI, iv_OSI, iv_SV):
        for fn in (cls.score, cls.make_decisions):           # first argument: `audios` (gmm_*) or `audio_list` (iv_*), as in the reference
            assert names(fn)[0] in ("audios", "audio_list") and set(names(fn)[1:]) == {"fs", "bits_per_sample", "debug", "n_jobs"}
    ref = "/root/reference"
    if os.path.isdir(ref):                                   # live check against the reference's own definitions
        import ast
        want = {}
        for mod in ("FAKEBOB", "gmm_ubm_CSI", "gmm_ubm_OSI", "gmm_ubm_SV", "ivector_PLDA_CSI", "ivector_PLDA_OSI", "ivector_PLDA_SV"):
            tree = ast.parse(open(os.path.join(ref, mod + ".py")).read())
            for node in ast.walk(tree):
                if isinstance(node, ast.ClassDef):
                    for fn in node.body:
                        if isinstance(fn, ast.FunctionDef) and fn.name in ("__init__", "score", "make_decisions", "attack", "estimate_threshold"):
                            want[(node.name, fn.name)] = [a.arg for a in fn.args.args if a.arg != "self"]
        ours = {"FakeBob": FakeBob, "gmm_CSI": gmm_CSI, "gmm_OSI": gmm_OSI, "gmm_SV": gmm_SV, "iv_CSI": iv_CSI, "iv_OSI": iv_OSI, "iv_SV": iv_SV}
        assert len(want) >= 20
        for (cname, fname), args in want.items():
            got = names(getattr(ours[cname], fname))
            assert got[:len(args)] == args, (cname, fname, args, got)


def test_library_contains_sm100a_tensor_core_and_tma_code():
    """The built library carries sm_100a SASS with tcgen05 MMAs (UTCHMMA), TMEM loads (LDTM), smem->TMEM copies (UTCCP) and
    TMA bulk copies (UBLKCP) -- i.e. the GMM hot path is the hand-written Blackwell kernel, not a generic fallback."""
    import shutil
    import subprocess
    from fakebob_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")
    elf = subprocess.run([cuobjdump, "--list-elf", _lib.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    assert "sm_100a" in elf, elf[:400]
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_Z15gmm_umma_kernelILb0ELb1ELi0EEv7GmmArgs", _lib.LIB_PATH],
                          capture_output=True, text=True, timeout=300).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTCCP", "UBLKCP", "SYNCS.PHASECHK"):
        assert mnemonic in sass, mnemonic


def test_library_ivector_kernels_use_tma_and_fp64_tensor_path():
    """The i-vector kernels of the built library: posteriors and the quadratic-term accumulation stage their operands with
    TMA bulk copies completing on mbarriers (UBLKCP, SYNCS), the Cholesky's trailing update runs on the FP64 tensor-core
    path (DMMA), the posteriors' quadratic forms use packed FFMA2."""
    import shutil
    import subprocess
    from fakebob_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the built library is not available")

    def sass(fn):
        return subprocess.run([cuobjdump, "-sass", "-fun", fn, _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    post = sass("_Z22fgmm_post_group_kernelPKfPKiS0_S0_S0_PKtS2_S2_S2_iiifPfS2_")
    for mnemonic in ("UBLKCP", "SYNCS.PHASECHK", "FFMA2", "LDS.64"):
        assert mnemonic in post, mnemonic
    quad = sass("_Z20ivec_quad_tma_kernelPKfPKdPKiiiiPdS4_")
    for mnemonic in ("UBLKCP", "SYNCS.PHASECHK", "DFMA"):
        assert mnemonic in quad, mnemonic
    solve = sass("_Z17ivec_solve_kernelPKdS0_iiiidPdPfPiPKi")
    assert solve.count("DMMA") >= 64 and "UCGABAR" in solve


def test_product_and_oracle_model_readers_agree(small_iv_tree):
    """The product's Kaldi file reader (fakebob_b200/kaldi_io.py) and the oracle's independent one (oracle/kaldi_files.py)
    parse every model file of the synthetic tree to the same arrays."""
    from oracle import kaldi_files as okf
    t = small_iv_tree
    pre = t["pre_model_dir"]
    a, b = kaldi_io.read_diag_gmm(t["ubm"]), okf.read_diag_gmm(t["ubm"])
    for k in ("weights", "gconsts", "means_invvars", "inv_vars"):
        assert np.array_equal(np.asarray(a[k]), b[k])
    a, b = kaldi_io.read_full_gmm(os.path.join(pre, "final.ubm")), okf.read_full_gmm(os.path.join(pre, "final.ubm"))
    for k in ("weights", "gconsts", "means_invcovars", "inv_covars"):
        assert np.array_equal(np.asarray(a[k]), b[k])
    a, b = kaldi_io.read_ivector_extractor(os.path.join(pre, "final.ie")), okf.read_ivector_extractor(os.path.join(pre, "final.ie"))
    assert np.array_equal(a["M"], b["M"]) and np.array_equal(a["sigma_inv"], b["sigma_inv"]) and a["prior_offset"] == b["prior_offset"]
    a, b = kaldi_io.read_plda(os.path.join(pre, "plda")), okf.read_plda(os.path.join(pre, "plda"))
    for k in ("mean", "transform", "psi"):
        assert np.array_equal(np.asarray(a[k]), b[k])
    assert np.array_equal(np.asarray(kaldi_io.read_vector(os.path.join(pre, "mean.vec"))), okf.read_vector(os.path.join(pre, "mean.vec")))
    assert np.array_equal(np.asarray(kaldi_io.read_matrix(os.path.join(pre, "transform.mat"))), okf.read_matrix(os.path.join(pre, "transform.mat")))


def test_bench_rooflines_bookkeeping():
    """bench.py's per-kernel roofline table on synthetic stage times: the GMM contraction's algorithmic FLOPs (SURVEY 8d:
    2 rows C 2D per model), the float64 kernels' fractions of the measured FMA rate, and the traffic table lookup."""
    import importlib
    import json
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    peaks = {"bf16_tflops": 1600.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6500.0}
    rows = 23205
    prof = {"gmm": (0.1 * 3, 3), "mfcc": (0.05 * 3, 3), "feats": (0.03 * 3, 3)}
    roof, ms = bench.rooflines("C2", prof, rows, 6, peaks)
    assert abs(ms["gmm"] - 0.1) < 1e-12
    flops = 2.0 * rows * 2048 * 144 * 6
    assert roof["gmm"]["algorithmic_flops_per_launch"] == flops
    assert abs(roof["gmm"]["achieved"] - flops / 0.1e-3 / 1e12) < 1e-6 and roof["gmm"]["peak"] == 1600.0
    assert abs(roof["gmm"]["frac"] * 1600.0 - roof["gmm"]["frac_of_sustained_peak"] * 1400.0) < 1e-6
    prof3 = {k: (v * 2, 2) for k, v in {"gmm": 0.13, "ivec_lin": 0.1, "ivec_quad": 0.16, "ivec_solve": 0.37, "fgmm_post": 0.23,
                                       "gselect": 0.05, "mfcc": 0.06, "feats": 0.03}.items()}
    roof3, _ = bench.rooflines("C3", prof3, rows, 1, peaks, active_frac=0.08)
    for k in ("ivec_lin", "ivec_quad", "ivec_solve"):
        f = roof3[k]["fp64"]
        assert f["peak_tfma_per_s"] == bench.FP64_TFMA_PEAK and 0.0 < f["frac"] < 1.0
    b = bench.CONFIGS["C3"]["S"] + 1
    assert roof3["ivec_solve"]["fp64"]["algorithmic_fma_per_launch"] == b * 400 ** 3 / 6.0
    assert roof3["ivec_quad"]["kernel"] == "ivec_quad_tma_kernel" and roof3["ivec_lin"]["kernel"] == "ivec_lin_tma_kernel"
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        table = json.load(f)
    for kernel in ("gmm_umma_kernel", "ivec_solve_kernel", "ivec_quad_tma_kernel", "ivec_lin_tma_kernel", "fgmm_post_group_kernel",
                   "gselect_kernel", "mfcc_kernel", "feats_kernel"):
        assert kernel in table and table[kernel]["dram_bytes_per_launch"] > 0
        assert bench.load_traffic(kernel)[0] == table[kernel]["dram_bytes_per_launch"]
