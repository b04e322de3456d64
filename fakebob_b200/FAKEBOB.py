"""FakeBob NES attack with the reference's API, executed on the B200.

Mirrors /root/reference FAKEBOB.py:
  FakeBob.__init__            :21-37   (same hyper-parameters; extra keyword-only knobs below)
  FakeBob.attack              :139-221 -> device loop (fb_nes_init / fb_nes_run / fb_nes_status)
  FakeBob.get_grad            :223-246 -> fb_nes_get_grad
  FakeBob.loss_fn             :248-299 -> same margin losses, evaluated on device inside the loop
  FakeBob.estimate_threshold  :39-137  -> host control flow, device scoring / gradient / update

Differences a caller can see, all deliberate:
  * ``rng``: "philox" (default) draws the antithetic Gaussian noise on the GPU with a counter-based
    generator whose seed is taken from numpy's global generator at construction, so ``np.random.seed``
    still makes runs reproducible; "numpy" reproduces the reference's stream exactly
    (``np.random.normal(size=(N, S//2))`` per iteration, FAKEBOB.py:234) at ~90 ms/iteration of host time.
  * ``model`` is normally one of this package's scorers (they carry the resident device models and the whole iteration
    then stays on the GPU).  Any other object with the reference's ``score()`` / ``make_decisions()`` interface
    (README.md:136, a black-box system) is attacked with the same device loop, except that the S+1 audios of each iteration
    are handed to ``model.score()`` on the host as int16-exact float audio (what the reference's own scorers quantise them to,
    gmm_ubm_OSI.py:83-85) and the scores are handed back; this mode uses the Philox noise stream.
  * per-iteration prints are emitted after each batch of ``iters_per_launch`` iterations.
"""
import os
import pickle
import time

import numpy as np

UNTARGETED = "untargeted"
TARGETED = "targeted"


class FakeBob(object):

    def __init__(self, task, attack_type, model, adver_thresh=0., epsilon=0.002, max_iter=1000,
                 max_lr=0.001, min_lr=1e-6, samples_per_draw=50, sigma=0.001, momentum=0.9,
                 plateau_length=5, plateau_drop=2., *, rng=None, seed=None, verbose=None, iters_per_launch=32):
        if not (hasattr(model, "score") and hasattr(model, "make_decisions")):
            raise TypeError("model must provide score() and make_decisions() (reference README.md:136)")
        self._external = not hasattr(model, "_engine")
        self._ext_engine = None
        self.task = task
        self.attack_type = attack_type
        self.model = model
        self.adver_thresh = adver_thresh
        self.epsilon = epsilon
        self.max_iter = max_iter
        self.max_lr = max_lr
        self.min_lr = min_lr
        self.samples_per_draw = samples_per_draw
        self.sigma = sigma
        self.momentum = momentum
        self.plateau_length = plateau_length
        self.plateau_drop = plateau_drop
        self.rng = rng or os.environ.get("FAKEBOB_RNG", "philox")
        if self.rng not in ("philox", "numpy"):
            raise ValueError("rng must be 'philox' or 'numpy'")
        if self._external and self.rng != "philox":
            raise ValueError("black-box models are attacked with the device noise generator (rng='philox')")
        # the Philox key comes from numpy's global generator (so np.random.seed still makes runs reproducible); in 'numpy'
        # mode nothing is drawn here, so the per-iteration np.random.normal stream is exactly the reference's
        if seed is not None:
            self.seed = int(seed)
        else:
            self.seed = int(np.random.randint(0, 2 ** 62)) if self.rng == "philox" else 0
        self.verbose = (os.environ.get("FAKEBOB_VERBOSE", "1") != "0") if verbose is None else verbose
        self.iters_per_launch = max(1, int(iters_per_launch))
        self.estimate_max_iters = 20000        # bound on the inner iterations of estimate_threshold (the reference loops forever)
        self.draws = 0                 # Philox draw counter == number of get_grad evaluations so far
        self.threshold = 0.
        self.true = None
        self.target = None

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _column(audio):
        audio = np.asarray(audio)
        if len(audio.shape) == 1:
            audio = audio[:, np.newaxis]
        elif audio.shape[0] == 1:
            audio = audio.T
        return audio

    def _label(self):
        if self.task == "CSI":
            return self.target if self.attack_type == TARGETED else self.true
        if self.task == "OSI" and self.attack_type == TARGETED:
            return self.target
        return None

    def _nes_init(self, audio, max_iter):
        m = self.model
        if self._external:
            from .engine import ExternalEngine
            if self._ext_engine is None:
                self._ext_engine = ExternalEngine()
            eng = self._ext_engine
            eng.nes_init(audio, self.task, self.attack_type, self._n_speakers(audio), self._label(), self.threshold,
                         self.adver_thresh, self.epsilon, max_iter, self.max_lr, self.min_lr, self.samples_per_draw,
                         self.sigma, self.momentum, self.plateau_length, self.plateau_drop, rng="philox", seed=self.seed,
                         draw_base=self.draws, external=True)
            return eng
        eng = m._engine
        zm = getattr(m, "z_norm_means", None) if self.task == "CSI" or m._fb_arch == "iv" else None
        zs = getattr(m, "z_norm_stds", None) if zm is not None else None
        eng.nes_init(audio, self.task, self.attack_type, getattr(m, "n_speakers", 1), self._label(),
                     self.threshold, self.adver_thresh, self.epsilon, max_iter, self.max_lr, self.min_lr,
                     self.samples_per_draw, self.sigma, self.momentum, self.plateau_length, self.plateau_drop,
                     rng=self.rng, seed=self.seed, draw_base=self.draws, z_means=zm, z_stds=zs)
        return eng

    def _n_speakers(self, audio):
        """Black-box models: the number of enrolled speakers is what score() returns per audio."""
        if self.task == "SV":
            return 1
        n = getattr(self.model, "n_speakers", None)
        if n is None:
            n = int(np.asarray(self.model.score(audio)).reshape(-1).shape[0])
        return n

    def _ext_scores(self, eng, kw):
        """One batch through a black-box scorer: (S+1, N) int16 from the device -> model.score((N, S+1) float) -> (S+1, K)."""
        wave = eng.nes_ext_perturb()
        audios = np.ascontiguousarray(wave.T.astype(np.float64) / 32768.0)
        sc = np.asarray(self.model.score(audios, **kw), dtype=np.float64)
        return sc.reshape(wave.shape[0], -1)

    def _host_noise(self, n):
        """The reference's draw (FAKEBOB.py:234), re-laid-out pair-major for the device."""
        pos = np.random.normal(size=(n, self.samples_per_draw // 2))
        return np.ascontiguousarray(pos.T)

    def _score_row(self, row):
        s = row[4:]
        return s[0] if self.task == "SV" else np.array(s)

    # ------------------------------------------------------------------------------------------
    def attack(self, audio, checkpoint_path, threshold=0., true=None, target=None, fs=16000,
               bits_per_sample=16, n_jobs=10, debug=False):
        audio = self._column(audio)
        self.threshold = threshold
        self.true = true
        self.target = target
        n = audio.shape[0]
        t_init = time.time()
        eng = self._nes_init(audio, self.max_iter)
        t_start = time.time()
        self.poll_times = [t_start - t_init]         # seconds: fb_nes_init, then one entry per polled batch of iterations
        done, stopped, printed = 0, 0, 0
        chunk = 1 if self.rng == "numpy" else self.iters_per_launch
        kw_score = dict(fs=fs, bits_per_sample=bits_per_sample, n_jobs=n_jobs, debug=debug)
        while not stopped and done < self.max_iter:
            k = min(chunk, self.max_iter - done)
            noise = None
            if self.rng == "numpy":
                noise = np.stack([self._host_noise(n) for _ in range(k)])
            t_poll = time.time()
            if self._external:
                eng.nes_ext_update(self._ext_scores(eng, kw_score))
            else:
                eng.nes_run(k, noise)
            done, stopped = eng.nes_status()
            self.poll_times.append(time.time() - t_poll)
            if self.verbose:
                rows = eng.nes_log(done)
                for it in range(printed, done):
                    r = rows[it]
                    print("--- iter %d, distance:%f, loss:%f, score: ---" % (it, r[0], r[1]), self._score_row(r))
                    if stopped and it == done - 1:
                        print("------ early stop at iter %d ---" % it)
                    else:
                        print("consumption time:%f, lr:%f" % ((time.time() - t_start) / max(done, 1), r[3]))
                printed = done
        elapsed = time.time() - t_start
        rows = eng.nes_log(done)
        self.draws += done
        per_iter = elapsed / max(done, 1)
        cp_global = []
        for it in range(done):
            r = rows[it]
            last_stop = stopped and it == done - 1
            cp_global.append([r[0], np.array([r[1]]), self._score_row(r), 0. if last_stop else per_iter])
        if checkpoint_path is not None:
            with open(checkpoint_path, "wb") as writer:
                pickle.dump(cp_global, writer, protocol=-1)
        self.log = rows
        self.iters_done = done
        self.elapsed = elapsed
        last_iter = done - 1
        success_flag = 1 if last_iter < self.max_iter - 1 else -1
        adver = eng.nes_adver()[:, np.newaxis]
        self.final_adver = adver
        adver = (adver * (2 ** (bits_per_sample - 1))).astype(np.int16)
        return adver, success_flag

    # ------------------------------------------------------------------------------------------
    def get_grad(self, audio, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        audio = self._column(audio)
        eng = self._nes_init(audio, 1)
        if self._external:
            eng.nes_ext_update(self._ext_scores(eng, dict(fs=fs, bits_per_sample=bits_per_sample, n_jobs=n_jobs, debug=debug)),
                               gradient_only=True)
            grad, losses, score0 = eng.nes_gest()
            self.draws += 1
            score = score0[0] if self.task == "SV" else score0
            return float(np.mean(losses[1:])), grad[:, np.newaxis], np.array([losses[0]]), score
        noise = self._host_noise(audio.shape[0]) if self.rng == "numpy" else None
        final_loss, grad, adver_loss, score0 = eng.nes_get_grad(noise)
        self.draws += 1
        score = score0[0] if self.task == "SV" else score0
        return final_loss, grad[:, np.newaxis], np.array([adver_loss]), score

    def loss_fn(self, audios, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        """Margin losses of FAKEBOB.py:248-299 for an explicit batch (host arithmetic on device scores)."""
        score = self.model.score(audios, fs=fs, bits_per_sample=bits_per_sample, n_jobs=n_jobs, debug=debug)
        if self.task in ("OSI", "CSI"):
            s2 = score if score.ndim == 2 else score[np.newaxis, :]
            if self.task == "OSI" and self.attack_type != TARGETED:
                loss = self.threshold + self.adver_thresh - np.max(s2, axis=1, keepdims=True)
            else:
                idx = self.target if self.attack_type == TARGETED else self.true
                other = np.max(np.delete(s2, idx, axis=1), axis=1, keepdims=True)
                own = s2[:, idx:idx + 1]
                if self.task == "OSI":
                    loss = np.maximum(other, self.threshold) + self.adver_thresh - own
                elif self.attack_type == TARGETED:
                    loss = other + self.adver_thresh - own
                else:
                    loss = own + self.adver_thresh - other
        else:
            loss = self.threshold + self.adver_thresh - np.asarray(score).reshape(-1)[:, np.newaxis]
        return loss, score

    # ------------------------------------------------------------------------------------------
    def estimate_threshold(self, audio, fs=16000, bits_per_sample=16, n_jobs=10, debug=False):
        if self.task == "CSI":
            print("--- Warning: no need to estimate threshold for CSI, quitting ---")
            return
        audio = self._column(audio)
        init_score = self.model.score(audio, fs=fs, bits_per_sample=bits_per_sample, n_jobs=n_jobs, debug=debug)
        if self.task == "OSI":
            init_score = np.max(init_score)
        self.delta = np.abs(init_score / 10)
        self.threshold = init_score + self.delta
        attack_type_backup = self.attack_type
        self.attack_type = UNTARGETED
        n = audio.shape[0]
        try:
            if self.rng == "philox" and not self._external:
                return self._estimate_threshold_device(audio)
            eng = self._nes_init(audio, 1)
            iter_outer, n_iters, times = 0, 0, 0.
            while True:
                if self.verbose:
                    print("----- iter_outer:%d, threshold:%f -----" % (iter_outer, self.threshold))
                eng.nes_set_threshold(self.threshold)
                iter_inner = 0
                lr = self.max_lr
                last_ls = []
                while True:
                    start = time.time()
                    adver = eng.nes_adver()[:, np.newaxis]
                    decision, score = self.model.make_decisions(adver, fs=fs, bits_per_sample=bits_per_sample,
                                                                n_jobs=n_jobs, debug=debug)
                    if self.verbose:
                        print("--- iter_inner:%d, dicision:%d, score: ---" % (iter_inner, decision), score)
                    if self.task == "OSI":
                        score = np.max(score)
                    if decision != -1:
                        if self.verbose:
                            print("--- return at iter_outer:%d, iter_inner:%d, return thresh:%f ---" % (iter_outer, iter_inner, score))
                            print("cost %d iters, %fs time" % (n_iters, times))
                        return score, n_iters, times
                    elif score >= self.threshold:
                        if self.verbose:
                            print("--- early stop at iter_inner:%d ---" % (iter_inner))
                        break
                    if self._external:
                        eng.nes_ext_update(self._ext_scores(eng, dict(fs=fs, bits_per_sample=bits_per_sample, n_jobs=n_jobs, debug=debug)),
                                           gradient_only=True)
                        loss = float(np.mean(eng.nes_gest()[1][1:]))
                    else:
                        loss, _, _, _ = eng.nes_get_grad(self._host_noise(n))
                    self.draws += 1
                    last_ls.append(loss)
                    last_ls = last_ls[-self.plateau_length:]
                    if last_ls[-1] > last_ls[0] and len(last_ls) == self.plateau_length:
                        if lr > self.min_lr:
                            lr = max(lr / self.plateau_drop, self.min_lr)
                        last_ls = []
                    eng.nes_apply_update(lr)
                    used_time = time.time() - start
                    if self.verbose:
                        print("consumption time:%f, lr:%f" % (used_time, lr))
                    n_iters += 1
                    times += used_time
                    iter_inner += 1
                self.threshold += self.delta
                iter_outer += 1
        finally:
            self.attack_type = attack_type_backup

    def _estimate_threshold_device(self, audio):
        """The search of FAKEBOB.py:76-137 with every inner iteration on the device (fb_nes_estimate_begin / fb_nes_continue):
        the make_decisions() score of the current adversarial audio is column 0 of the iteration's own batch, so no second
        scoring call and no host round trip per inner iteration.  Same control flow, prints and return value."""
        eng = self._nes_init(audio, self.estimate_max_iters)
        eng.nes_estimate_begin(self.model.threshold)
        iter_outer, n_iters, times, rows_seen = 0, 0, 0., 0
        chunk = self.iters_per_launch
        while True:
            if self.verbose:
                print("----- iter_outer:%d, threshold:%f -----" % (iter_outer, self.threshold))
            eng.nes_continue(self.threshold)
            start = time.time()
            stopped = 0
            while not stopped:
                eng.nes_run(chunk)
                done, stopped = eng.nes_status()
                if not stopped and done >= self.estimate_max_iters:
                    raise RuntimeError("estimate_threshold: no decision after %d inner iterations" % self.estimate_max_iters)
            rows = eng.nes_log(done)
            elapsed = time.time() - start
            n_new = done - rows_seen                    # inner iterations of this outer iteration, the last one is the stop test
            per_iter = elapsed / max(n_new, 1)
            for i in range(n_new):
                r = rows[rows_seen + i]
                score = self._score_row(r)
                last = i == n_new - 1
                if self.verbose:
                    print("--- iter_inner:%d, dicision:%d, score: ---" % (i, self._decision(score) if (last and stopped == 2) else -1), score)
                    if not last:
                        print("consumption time:%f, lr:%f" % (per_iter, r[3]))
            n_iters += n_new - 1
            times += per_iter * (n_new - 1)
            self.draws += n_new - 1
            rows_seen = done
            final = rows[done - 1]
            score = self._score_row(final)
            if self.task == "OSI":
                score = np.max(score)
            if stopped == 2:
                if self.verbose:
                    print("--- return at iter_outer:%d, iter_inner:%d, return thresh:%f ---" % (iter_outer, n_new - 1, score))
                    print("cost %d iters, %fs time" % (n_iters, times))
                return score, n_iters, times
            if self.verbose:
                print("--- early stop at iter_inner:%d ---" % (n_new - 1))
            self.threshold += self.delta
            iter_outer += 1

    def _decision(self, score):
        """make_decisions() of the scorer for one audio, from its scores (gmm_ubm_OSI.py:93-112, gmm_ubm_SV.py:81-92)."""
        if self.task == "OSI":
            return int(np.argmax(score)) if np.max(score) >= self.model.threshold else -1
        return 1 if score >= self.model.threshold else -1
