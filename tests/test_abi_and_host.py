"""CPU tests: the C-ABI library loads and exports every symbol include/spectral_b200.h declares, the product
path fails loudly without a GPU (no CPU fallback, nothing routes through oracle/), and the host-side logic
(configuration adapter, scale constants, phase draws, utterance sharding) is right."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sb():
    import __graft_entry__ as g
    g.build()                         # nvcc cross-compiles for sm_100a without a GPU
    import transtacos_retunegan_b200 as sb
    return sb


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "spectral_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z_0-9]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(sb):
    lib = ctypes.CDLL(sb._lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/spectral_b200.h but not exported"
    assert set(names) == set(sb._lib.SIGNATURES), set(names) ^ set(sb._lib.SIGNATURES)
    assert b"sm_100a" in sb._lib.load().sb200_version()


def test_library_is_built_for_sm100a_only(sb):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", sb._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_path_fails_loudly_without_gpu(sb):
    y = np.zeros(8192, np.float32)
    for fn in (lambda: sb.transtacos_audio.get_specs(y), lambda: sb.retunegan_audio.get_mag(y),
               lambda: sb.retunegan_audio.inv_mag(np.zeros((1025, 10), np.float32)),
               lambda: sb.multi_stft_loss(torch.zeros(1, 4096), torch.zeros(1, 4096), ret_loss=True)):
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()
    # the C ABI itself reports the missing device instead of computing anything on the host
    cfg = sb._lib.Config(22050, 2048, 1024, 256, 80, 125.0, 7600.0, 0, 0)
    h = ctypes.c_void_p()
    rc = sb._lib.load().sb200_plan_create(ctypes.byref(cfg), ctypes.byref(h))
    assert rc == -2 and b"no CUDA device" in sb._lib.load().sb200_last_error_string()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "transtacos-retunegan_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle|import_module\(.oracle|/oracle/|oracle\.", re.M)
    n = 0
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h")):
                n += 1
                assert not pat.search(open(os.path.join(dirpath, f)).read()), f"{f} references oracle/"
    assert n >= 12


def test_config_adapter_and_validation(sb):
    hp = types.SimpleNamespace(sample_rate=16000, n_fft=1024, win_length=512, hop_length=128, n_mel=40, n_freq=513,
                               fmin=50, fmax=7000, gl_iters=7, multi_stft_params=[(1024, 512, 128)], unrelated=1)
    c = sb.SpectralConfig.from_hparam(hp)
    assert (c.sample_rate, c.n_fft, c.n_mel, c.gl_iters, c.multi_stft_params) == (16000, 1024, 40, 7, ((1024, 512, 128),))
    assert c.plan_key() == (16000, 1024, 512, 128, 40, 50.0, 7000.0, 0, 0)
    assert sb.TRANSTACOS.gl_iters == 30 and sb.RETUNEGAN.gl_iters == 4 and sb.RETUNEGAN.gl_momentum == 0.7
    for bad in (dict(window_fn="kaiser"), dict(mel_scale="bark"), dict(n_freq=1024), dict(fmax=11025)):
        with pytest.raises(ValueError):
            sb.SpectralConfig(**bad)
    assert sb.PI == 3.14159265358979


def test_scale_constants_match_the_reference_formulas(sb):
    from oracle import spectral_oracle as O
    s = sb.transtacos_audio.db_norm_scale(sb.TRANSTACOS)
    x = np.array([1e-7, 1e-5, 3e-3, 0.5, 12.0])
    ours = s.a * np.log2(np.maximum(s.floor, x)) + s.b
    np.testing.assert_allclose(ours, O._normalize(O._amp_to_db(x) - 20), atol=1e-6)   # float32 struct fields
    ln = sb.retunegan_audio.ln_scale(True)
    np.testing.assert_allclose(ln.a * np.log2(np.maximum(ln.floor, x)) + ln.b, np.log(x.clip(min=1e-5)), rtol=1e-6)
    assert sb.retunegan_audio.ln_scale(False).floor == 0.0


def test_phase_draws_match_the_reference_rng(sb):
    np.random.seed(114514)
    a = sb.transtacos_audio.draw_phase(1025, 7)          # global RNG, like transtacos/audio.py:134
    np.random.seed(114514)
    np.testing.assert_array_equal(a, np.random.rand(1025, 7))
    b = sb.transtacos_audio.draw_phase(1025, 7, seed=114514)   # fresh RandomState, like librosa.griffinlim
    np.testing.assert_array_equal(b, np.random.RandomState(114514).rand(1025, 7))


def test_host_helpers_match_oracle(sb):
    from oracle import spectral_oracle as O
    ta = sb.transtacos_audio
    S = np.random.RandomState(0).uniform(-5.6, 4.0, (1024, 9))
    np.testing.assert_allclose(ta.spec_to_natural_scale(S), O.tt_spec_to_natural_scale(S), rtol=1e-12)
    np.testing.assert_allclose(ta.fix_zero_DC(ta.spec_to_natural_scale(S)), O.tt_fix_zero_DC(O.tt_spec_to_natural_scale(S)))
    assert ta.fix_zero_DC(np.ones((1025, 3))).shape == (1025, 3)
    np.testing.assert_allclose(ta._normalize(ta._amp_to_db(np.array([1e-9, 0.3]))), O._normalize(O._amp_to_db(np.array([1e-9, 0.3]))))
    assert len(ta.align_wav(np.zeros(1000))) == 1024 and len(ta.align_wav(np.zeros(1024))) == 1024


def test_shard_utterances_balanced_and_deterministic(sb):
    rs = np.random.RandomState(114514)
    T = np.clip(np.round(rs.normal(307, 100, 10000)), 101, 524).astype(int)
    lens = 256 * T - 1
    shards = sb.sharding.shard_utterances(lens, 8)
    assert sorted(i for s in shards for i in s) == list(range(10000))
    loads = np.array([lens[s].sum() for s in shards])
    assert loads.max() / loads.min() < 1.001
    assert shards == sb.sharding.shard_utterances(lens, 8)
    assert sb.sharding.shard_utterances([5, 3, 8], 1) == [[2, 0, 1]]
    assert [len(s) for s in sb.sharding.shard_utterances([1, 1], 4)] == [1, 1, 0, 0]
    with pytest.raises(ValueError):
        sb.sharding.shard_utterances([1], 0)


def test_numa_binding_is_a_no_op_without_nvml():
    """sharding.bind_to_gpu_numa never raises: without a GPU / NVML it reports False and leaves the affinity alone."""
    import os
    import transtacos_retunegan_b200 as sb
    before = os.sched_getaffinity(0)
    ok = sb.sharding.bind_to_gpu_numa(0)
    assert ok in (True, False)
    if not ok:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_peer_reduce_descriptor_layout():
    """sb200_peer_reduce (include/spectral_b200.h): 8 buffer pointers, rank, world, the output pointer -- the ctypes mirror must
    have the C layout, and the exchange buffer is the 256-byte PeerBuf of mstft.cuh.  No GPU needed."""
    import ctypes as C
    import transtacos_retunegan_b200 as sb
    L = sb._lib
    assert L.MAX_PEERS == 8 and C.sizeof(L.PeerReduce) == 8 * 8 + 4 + 4 + 8
    assert L.PeerReduce.rank.offset == 64 and L.PeerReduce.world.offset == 68 and L.PeerReduce.loss_global.offset == 72
    assert L.load().sb200_peer_buffer_bytes() == 256
    red = sb.loss.PeerLossReducer(1, 2, peers=[4096, 8192])          # single-process form: raw pointers, nothing is allocated
    t = __import__("torch").zeros(())
    d = red.descriptor(t)
    assert (d.rank, d.world, d.peer[0], d.peer[1], d.peer[2], d.loss_global) == (1, 2, 4096, 8192, None, t.data_ptr())
    with pytest.raises(ValueError):
        sb.loss.PeerLossReducer(0, 9, peers=[0] * 9)


def test_fast_atan2_polynomial_of_the_spec_stack_kernels():
    """mstft.cuh fast_atan2 (phase channel of the spec stacks / get_stft_torch): the degree-7 minimax coefficients are read from the
    source and the kernel's arithmetic is replayed in float32 numpy over all four quadrants, the axes and a range of magnitudes:
    within 4e-7 rad of arctan2 (the header says ~3e-7), and a zero spectrum has phase 0 whatever the sign of its zeros."""
    import re
    src = open(os.path.join(ROOT, "transtacos-retunegan_b200", "csrc", "mstft.cuh")).read()
    body = src[src.index("float fast_atan2(float y, float x)"):]
    body = body[:body.index("return copysignf")]
    coef = [np.float32(c) for c in re.findall(r"(-?\d\.\d+(?:e-?\d+)?)f\)?;", body) if abs(float(c)) < 1.01]
    assert len(coef) == 8 and abs(coef[-1] - 1) < 1e-6, coef          # r = c7; r = fma(r, s, c6) ... ; last is ~1
    rs = np.random.RandomState(0)
    x = np.concatenate([rs.randn(200000) * 10 ** rs.uniform(-6, 3, 200000), [0, 0, 1, -1, 0, 0, -0.0, 3, -3]]).astype(np.float32)
    y = np.concatenate([rs.randn(200000) * 10 ** rs.uniform(-6, 3, 200000), [1, -1, 0, 0, 0, -0.0, 0, 3, 3]]).astype(np.float32)
    ax, ay = np.abs(x), np.abs(y)
    mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
    a = (mn / np.maximum(mx, np.float32(1e-37))).astype(np.float32)
    s = (a * a).astype(np.float32)
    r = np.full_like(a, coef[0])
    for c in coef[1:]:
        r = (r * s + c).astype(np.float32)
    r = (r * a).astype(np.float32)
    r = np.where(ay > ax, np.float32(1.57079632679489662) - r, r).astype(np.float32)
    r = np.where(x < 0, np.float32(3.14159265358979324) - r, r).astype(np.float32)
    r = np.copysign(r, y)
    ref = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    keep = ~((x == 0) & (y == 0))
    assert np.abs(r[keep] - ref[keep]).max() < 4e-7, np.abs(r[keep] - ref[keep]).max()
    assert np.all(r[~keep] == 0)                                           # digital silence: angle(0) = 0 like torch.angle
