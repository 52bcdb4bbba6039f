"""TEST INFRASTRUCTURE -- imports the UNMODIFIED reference package from /root/reference.

Only usable inside the build container (``/root/reference`` does not exist on the
GPU box).  It is used by ``oracle/make_golden.py`` to produce the committed golden
vectors under ``tests/golden/`` and by ``tests/test_oracle_vs_reference.py`` (skipped
when the reference tree is absent).

The reference (``asr_deepspeech`` v0.4.10) imports six packages that are not
installed here and carry *no arithmetic on the hot path*: gnutools, sakura,
ascii_graph, librosa, soundfile, Levenshtein.  They are
replaced by inert in-process stubs so that ``asr_deepspeech.modules`` /
``asr_deepspeech.functional`` import exactly as shipped.  ``librosa.stft`` is the
one stub that *would* carry arithmetic; it raises, so nothing here can silently
pretend to be librosa (the STFT oracle is a restatement, see oracle/explicit.py).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ASR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "asr_deepspeech"))


def _stub(modname: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(modname)
    mod.__dict__.update(attrs)
    sys.modules[modname] = mod
    return mod


def _install_stubs() -> None:
    import yaml

    class _NS(dict):
        __getattr__ = dict.get

    def load_config(path):  # gnutools.fs.load_config: YAML -> attribute namespace
        with open(path) as f:
            raw = yaml.safe_load(f)

        def wrap(x):
            return _NS({k: wrap(v) for k, v in x.items()}) if isinstance(x, dict) else x

        return wrap(raw)

    fs = _stub(
        "gnutools.fs",
        load_config=load_config,
        parent=os.path.dirname,
        listfiles=lambda root, patterns=None: [],
        name=lambda p: os.path.splitext(os.path.basename(p))[0],
    )
    _stub("gnutools", fs=fs)
    _stub("gnutools.concurrent", ProcessPoolExecutorBar=object)
    _stub("gnutools.tests", test_imports=lambda *a, **k: None)

    class SakuraTrainer:  # sakura.ml.SakuraTrainer: attribute bag, as used by the reference
        def __init__(self, model, optimizer, scheduler, metrics, epochs, model_path,
                     checkpoint_path, device, device_test):
            self._model, self._optimizer, self._scheduler = model, optimizer, scheduler
            self._metrics, self._model_path = metrics, model_path
            self._device, self._device_test = device, device_test

    _stub("sakura")
    _stub("sakura.ml", SakuraTrainer=SakuraTrainer, AsyncTrainer=lambda trainer: trainer)
    _stub("sakura.functional", asr_metrics=None)

    class Pyasciigraph:
        def graph(self, title, data):
            return [title] + [f"{k}: {v}" for k, v in data]

    _stub("ascii_graph", Pyasciigraph=Pyasciigraph)

    def _no_librosa(*a, **k):
        raise RuntimeError("librosa is not installed; the STFT oracle is oracle.explicit.spectrogram")

    _stub("librosa", stft=_no_librosa, magphase=_no_librosa,
          util=types.SimpleNamespace(find_files=lambda *a, **k: []))
    _stub("soundfile", read=_no_librosa, write=_no_librosa, info=_no_librosa)

    def _lev(a, b):  # Levenshtein.distance (host metric, out of scope; kept exact)
        prev = list(range(len(b) + 1))
        for i, ca in enumerate(a, 1):
            cur = [i]
            for j, cb in enumerate(b, 1):
                cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
            prev = cur
        return prev[-1]

    _stub("Levenshtein", distance=_lev)


_loaded = None


def load_reference():
    """Return the reference ``asr_deepspeech`` package (imported unmodified)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch

    rng = torch.get_rng_state()  # asr_deepspeech/vars.py:13 reseeds the global RNG at import
    _install_stubs()
    os.environ.setdefault("ZAK_ASR_CONFIG", os.path.join(REFERENCE_ROOT, "asr_deepspeech", "config.yml"))
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        pkg = importlib.import_module("asr_deepspeech")
        importlib.import_module("asr_deepspeech.modules.blocks")
        importlib.import_module("asr_deepspeech.modules.deepspeech")
        importlib.import_module("asr_deepspeech.decoders.greedy_decoder")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        torch.set_rng_state(rng)
    _loaded = pkg
    return pkg
