"""CPU oracle for the FAKEBOB hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``fakebob_b200``) never does; it fails loudly when its CUDA library is missing.

What is restated here
---------------------
* ``oracle.nes``          -- FakeBob.attack / get_grad / loss_fn / estimate_threshold
                             (reference ``FAKEBOB.py:39-299``).  PINNED: checked against
                             the reference's own ``FAKEBOB.py`` imported from
                             ``/root/reference`` (``tests/golden/make_golden.py``), and the
                             resulting trajectories are committed under ``tests/golden/``.
* ``oracle.kaldi_feats``  -- Kaldi compute-mfcc-feats / compute-vad / add-deltas /
                             apply-cmvn-sliding / select-voiced-frames as invoked at
                             ``gmm_ubm_kaldiHelper.py:138,158,195-198``.
* ``oracle.diag_gmm``     -- Kaldi gmm-global-get-frame-likes --average=true
                             (``gmm_ubm_kaldiHelper.py:206-208``) and MAP mean adaptation
                             (``build_spk_models.py:202-219``, ``gmm-global-est-map.cc``).
* ``oracle.ivector``      -- gmm-gselect / fgmm-global-gselect-to-post / ivector-extract
                             (``ivector_PLDA_kaldiHelper.py:202``) and the PLDA back-end
                             (``ivector_PLDA_kaldiHelper.py:262-271``).
* ``oracle.scorers``      -- the six task wrappers' score()/make_decisions() rules
                             (``gmm_ubm_{CSI,OSI,SV}.py``, ``ivector_PLDA_{CSI,OSI,SV}.py``).

PARITY UNPINNED for the Kaldi arithmetic: the reference executes un-vendored,
un-pinned upstream Kaldi binaries (kaldi-asr/kaldi HEAD, ``docker/Dockerfile:24``)
which are absent from /root/reference and from this image, and the reference ships
no golden vectors (``test.py`` only prints).  The Kaldi restatement follows the
published upstream algorithm and is guarded against
transcription errors by independent implementations available in-container:
``torchaudio.compliance.kaldi.mfcc`` (MFCC), ``sklearn.mixture.GaussianMixture``
(diag / full GMM log-likelihoods and posteriors) and closed-form float64 algebra
(i-vector solve, PLDA LLR).  See tests/test_oracle_*.py.
"""
