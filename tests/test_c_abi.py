"""CPU: the C-ABI shared library builds/loads and exports exactly what include/asr_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os

import pytest


def test_library_exports_every_header_symbol():
    from asr_b200 import _lib

    assert os.path.exists(_lib.LIBRARY), "run __graft_entry__.build() first"
    names = set(_lib.PROTOTYPES)
    expected = {"asrb_version", "asrb_strerror", "asrb_set_debug_flags", "asrb_gemm_tn", "asrb_rnn_fwd", "asrb_rnn_fwd_sum",
                "asrb_rnn_bwd", "asrb_conv2d_mask_fwd", "asrb_bn_act_mask_fwd", "asrb_bn_rows_fwd",
                "asrb_log_softmax_fwd", "asrb_ctc_fwd", "asrb_ctc_bwd", "asrb_spectrogram"}
    assert expected <= names
    dll = ctypes.CDLL(_lib.LIBRARY)
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/asr_b200.h but not exported"
    assert _lib.query("asrb_version") >= 100


def test_error_contract_without_a_gpu():
    """negative status = bad argument / unsupported shape, decoded by asrb_strerror; nothing is launched"""
    from asr_b200 import _lib

    assert "bad argument" in _lib.strerror(-1)
    assert "unsupported" in _lib.strerror(-3)
    with pytest.raises(ValueError):           # NULL pointers are rejected before any CUDA call
        _lib.call("asrb_gemm_tn", None, 4, None, 4, None, 4, None, 1, 1, 4, 0, None)
    nj, P = ctypes.c_int(), ctypes.c_int()
    for bf16 in (0, 1):
        _lib.call("asrb_rnn_plan", 0, 800, 64, bf16, ctypes.byref(nj), ctypes.byref(P), None, None)
        assert 2 * P.value <= 148 and nj.value * P.value >= 800  # one CTA per SM, both directions co-resident
    with pytest.raises(ValueError):           # hidden size whose rows are not 16-byte multiples (TMA)
        _lib.call("asrb_rnn_plan", 0, 801, 64, 0, None, None, None, None)
    with pytest.raises(ValueError):           # fp32/tf32 LSTM-1024 slices do not fit shared memory: bf16 only
        _lib.call("asrb_rnn_plan", 1, 1024, 128, 0, None, None, None, None)
    _lib.call("asrb_rnn_plan", 1, 1024, 128, 1, ctypes.byref(nj), ctypes.byref(P), None, None)
    assert 2 * P.value <= 148
    with pytest.raises(ValueError):           # the newer entry points keep the contract: NULL / non-positive sizes
        _lib.call("asrb_lookahead_fwd", None, None, None, 4, 2, 8, 3, 0, 0.0, 0.0, None)
    with pytest.raises(ValueError):
        _lib.call("asrb_lookahead_bwd", None, None, None, None, None, None, 4, 2, 8, 0, 0, 0.0, 0.0, None)
    with pytest.raises(ValueError):
        _lib.call("asrb_adamw_step", None, None, None, None, 16, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, None, None)
    with pytest.raises(ValueError):           # row-grouped conv entry points: NULL pointers, rows per item other than 2 / 4
        _lib.call("asrb_conv32_pack_rows", None, None, 21, 11, 2, 4, 0, None)
    with pytest.raises(ValueError):
        _lib.call("asrb_conv32_fwd_rows", None, None, None, None, None, 1, 81, 100, 41, 100, 21, 11, 2, 10, 5, 3, None)
    with pytest.raises(ValueError):
        _lib.call("asrb_conv32_bwd_data_rows", None, None, None, 1, 81, 100, 41, 100, 21, 11, 2, 10, 5, 4, None)
    assert _lib.query("asrb_gemm_cta_limit", -1) == 0 and _lib.query("asrb_debug_gemm_tma_store", -1) == 1
    assert _lib.query("asrb_debug_conv_wgrad_bf16", -1) == 1
    assert _lib.query("asrb_conv32_bwd_weight_workspace_bytes", 64, 81, 501, 41, 501) >= 8 * 64 * 32 * 81 * 504 * 2
    ws = _lib.query("asrb_ctc_workspace_bytes", 2000, 256, 200)  # alpha and gathered log-prob rows [N,T,row stride >= 2U+1] f32 + target offsets
    assert 2 * 2000 * 256 * 401 * 4 <= ws <= 2 * 2000 * 256 * 416 * 4 + 2 * 256 * 4 + 64


def test_sass_contains_tcgen05_and_tma():
    """The product GEMM / recurrence really are Blackwell-native: UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld),
    UTMALDG (TMA) in the SASS of the shipped library."""
    import shutil
    import subprocess

    from asr_b200 import _lib

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIBRARY], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTMAREDG"):   # + TMA store / reduce (GEMM epilogue)
        assert mnemonic in sass, mnemonic
