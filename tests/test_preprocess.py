"""GPU-sharded preprocessing driver and its on-disk formats against the reference's
transtacos/preprocess.py:16-41 and transtacos/datasets/databaker.py:25-160.

CPU part: label parsing, the 2-sigma filter / corpus statistics, the metadata files and the .npy formats (host logic,
checked against a line-by-line restatement of the reference and against what its consumer, transtacos/data.py:153-161,
does with the files).  GPU part: a small synthetic DataBaker-shaped corpus end to end against the oracle.
"""
import os
import random
import types

import numpy as np
import pytest
import torch

from conftest import rel_fro
from oracle import spectral_oracle as O


def _mod():
    import importlib
    return importlib.import_module("transtacos_retunegan_b200").preprocess


LABELS = ("000001\t卡尔普#2陪外孙#1玩滑梯#4。\n\tka2 er2 pu3 pei2 wai4 sun1 wan2 hua2 ti1\n"
          "000002\t假语村言#2别再#1拥抱我#4。\n\tjia2 yu3 cun1 yan2 bie2 zai4 yong1 bao4 wo3\n"
          "000003\t宝马#1配挂#1跛骡鞍#3，貂蝉#1怨枕#2董翁榻#4。\n\tbao2 ma3 pei4 gua4 bo3 luo2 an1 diao1 chan2 yuan4 zhen3 dong3 weng1 ta4\n")


def _ref_parse(text):
    """datasets/databaker.py:125-160 restated on a string."""
    import re
    punct = re.compile(r'，|。|、|：|；|？|！|（|）|“|”|…|—')
    r, lines = {}, text.split("\n")
    for i in range(0, len(lines) - 1, 2):
        if not lines[i].strip():
            break
        name, kanji = lines[i].strip().split("\t")
        pinyin = lines[i + 1].strip().lower()
        kanji = punct.sub('', kanji)
        pr = []
        for k in kanji:
            if k == '#':
                continue
            if k.isdigit():
                if pr: pr[-1] = k
                else: pr.append(k)
            else:
                pr.append('0')
        r[name] = (pinyin, ''.join(pr))
    return r


def test_parse_label_file(tmp_path):
    P = _mod()
    fp = tmp_path / "labels.txt"
    fp.write_text(LABELS, encoding="utf-8")
    got = P.parse_label_file(str(fp))
    assert got == _ref_parse(LABELS)
    assert got["000001"] == ("ka2 er2 pu3 pei2 wai4 sun1 wan2 hua2 ti1", "002001004")
    for text, prds in got.values():
        assert len(text.split(" ")) == len(prds)


def _fake_meta(n, seed=0):
    rs = np.random.RandomState(seed)
    out = []
    for i in range(n):
        lt = int(rs.randint(5, 30)) if i else 200           # one text-length outlier
        T = int(rs.randint(100, 500)) if i != 1 else 5000   # one audio-length outlier
        st = {'max_mel': rs.rand(), 'min_mel': -rs.rand(), 'max_mag': rs.rand(), 'min_mag': -rs.rand(),
              'max_c0': rs.rand(), 'min_c0': rs.rand() * 1e-3}
        out.append((f"{i:06d}", "0" * lt, " ".join(["a1"] * lt), lt, T * 256, T, st))
    return out


def test_filter_aggregate_and_metadata_files(tmp_path):
    P = _mod()
    meta = _fake_meta(60)
    kept, stats = P.filter_and_aggregate(list(meta) + [None], 22050)
    # restatement of datasets/databaker.py:39-88
    tl = np.asarray([m[-4] for m in meta]); al = np.asarray([m[-2] for m in meta])
    ok = [m for m in meta if tl.mean() - 2 * tl.std() <= m[-4] <= tl.mean() + 2 * tl.std()
          and al.mean() - 2 * al.std() <= m[-2] <= al.mean() + 2 * al.std()]
    assert [m[:3] for m in ok] == kept and "000000" not in [m[0] for m in kept] and "000001" not in [m[0] for m in kept]
    assert stats['total_examples'] == len(ok)
    assert stats['max_len_spec'] == max(m[-2] for m in ok) and stats['min_len_txt'] == min(m[-4] for m in ok)
    assert np.isclose(stats['total_hours'], sum(m[-3] for m in ok) / 22050 / 3600)
    assert stats['max_mel'] == max(m[-1]['max_mel'] for m in ok) and stats['min_c0'] == min(m[-1]['min_c0'] for m in ok)
    # metadata files (transtacos/preprocess.py:16-41)
    P.write_metadata(list(kept), stats, "/data/DataBaker/Wave", str(tmp_path), "pp", shuffle=True, split_ratio=0.05, seed=114514)
    ref = list(kept)
    random.seed(114514)
    random.shuffle(ref)
    cp = int(len(ref) * 0.05)
    rd = lambda fn: (tmp_path / "pp" / fn).read_text(encoding="utf-8")
    assert rd("test.txt") == "".join("|".join(str(x) for x in m) + "\n" for m in ref[:cp])
    assert rd("train.txt") == "".join("|".join(str(x) for x in m) + "\n" for m in ref[cp:])
    assert rd("wav_path.txt") == "/data/DataBaker/Wave"
    assert rd("stats.txt").splitlines()[0] == f"total_examples\t{len(ok)}"
    assert len(rd("stats.txt").splitlines()) == len(stats)


def test_npy_formats_match_the_consumer_contract(tmp_path):
    P = _mod()
    T, F, M = 37, 1025, 80
    rs = np.random.RandomState(1)
    mag_fm = torch.from_numpy(rs.randn(T, F).astype(np.float32))
    mel_fm = torch.from_numpy(rs.randn(T, M).astype(np.float32))
    c0 = torch.from_numpy(rs.rand(T).astype(np.float32))
    st = P.save_features(str(tmp_path), "000007", mag_fm, mel_fm, c0, f0=np.arange(T))
    mel = np.load(tmp_path / "mel-000007.npy"); mag = np.load(tmp_path / "mag-000007.npy")
    assert mel.shape == (M, T) and mag.shape == (F, T) and mel.dtype == np.float64 and mag.dtype == np.float64
    assert mag.flags.f_contiguous and mel.flags.f_contiguous          # what np.abs(librosa.stft(...)) arithmetic produces
    # the consumer (transtacos/data.py:153-161) transposes and drops the DC row
    np.testing.assert_array_equal(mag.T[:, 1:], mag_fm.numpy()[:, 1:].astype(np.float64))
    np.testing.assert_array_equal(mel.T, mel_fm.numpy().astype(np.float64))
    assert np.load(tmp_path / "c0-000007.npy").dtype == np.float32 and np.load(tmp_path / "f0-000007.npy").dtype == np.float32
    assert st['max_mag'] == mag.max() and st['min_mel'] == mel.min() and st['max_f0'] == T - 1
    with open(tmp_path / "mag-000007.npy", "rb") as fh:
        assert b"'fortran_order': True" in fh.read(128)


def _amp(S):   # inverse of _normalize / _amp_to_db (transtacos/audio.py:177-196)
    return 10.0 ** ((((np.asarray(S, np.float64) + 4) * 100 / 8) - 100 + 20) / 20)


@pytest.mark.gpu
def test_preprocess_small_corpus_end_to_end(tmp_path):
    import transtacos_retunegan_b200 as sb
    from scipy.io import wavfile
    assert torch.cuda.is_available()
    P = sb.preprocess
    base = tmp_path
    (base / "DataBaker" / "Wave").mkdir(parents=True)
    (base / "DataBaker" / "ProsodyLabeling").mkdir(parents=True)
    (base / "DataBaker" / "ProsodyLabeling" / "000001-010000.txt").write_text(LABELS, encoding="utf-8")
    wavs = {}
    for i, (name, L) in enumerate([("000001", 30000), ("000002", 41234), ("000003", 25600)]):
        y = (1e-4 * O.synth_noise(L + 8000, i)).astype(np.float32)
        y[4000:4000 + L] += O.synth_speechlike(L, 100 + i)
        if i == 1:
            wavfile.write(base / "DataBaker" / "Wave" / f"{name}.wav", 22050, np.round(y * 32768).astype(np.int16))
            y = np.round(y * 32768).astype(np.int16).astype(np.float32) / 32768
        else:
            wavfile.write(base / "DataBaker" / "Wave" / f"{name}.wav", 22050, y)
        wavs[name] = y
    args = types.SimpleNamespace(base_dir=str(base), out_dir="preprocessed")
    old = P.DROPOUT_2SIGMA
    P.DROPOUT_2SIGMA = False        # three utterances: keep them all
    try:
        metadata, stats, wav_dp = P.preprocess(args)
    finally:
        P.DROPOUT_2SIGMA = old
    assert wav_dp == str(base / "DataBaker" / "Wave") and [m[0] for m in metadata] == ["000001", "000002", "000003"]
    labels = _ref_parse(LABELS)
    tot = 0
    for name, prds, text in metadata:
        assert (text, prds) == labels[name]
        y = O.tt_trim_silence(wavs[name])                       # datasets/databaker.py:97-103 restated on the oracle
        d = len(y) % 256
        y = np.pad(y, (0, 256 - d)) if d else y
        So, Mo = O.tt_get_specs(y[:-1])
        mag = np.load(base / "preprocessed" / f"mag-{name}.npy"); mel = np.load(base / "preprocessed" / f"mel-{name}.npy")
        c0 = np.load(base / "preprocessed" / f"c0-{name}.npy")
        assert mag.shape == So.shape == (1025, len(y) // 256) and mel.shape == Mo.shape and mag.flags.f_contiguous
        assert rel_fro(_amp(mag), _amp(So)) < 1e-4 and rel_fro(_amp(mel), _amp(Mo)) < 1e-4
        np.testing.assert_allclose(c0, O.tt_get_c0(y[:-1]), rtol=1e-5, atol=1e-9)
        f0 = np.load(base / "preprocessed" / f"f0-{name}.npy")
        assert f0.shape == c0.shape and f0.dtype == np.float32
        assert (np.abs(f0 / O.tt_get_f0(y[:-1]) - 1) < 1e-4).mean() >= 0.99
        tot += len(y)
    assert stats['total_examples'] == 3 and np.isclose(stats['total_hours'], tot / 22050 / 3600)
    assert stats['min_mag'] >= -5.6 - 1e-3                       # floor of the normalised dB scale (stats/DataBaker.stats:13)
    P.write_metadata(metadata, stats, wav_dp, str(base), "preprocessed")
    assert sorted(os.listdir(base / "preprocessed"))[-4:] == ["stats.txt", "test.txt", "train.txt", "wav_path.txt"]


@pytest.mark.gpu
def test_too_short_clip_is_skipped_not_fatal(tmp_path):
    """One clip that is too short for the centred window (after trimming) must not abort the batch: the reference handles clips
    one by one (datasets/databaker.py:91-122), so the others are still written and the short one comes back as None."""
    import transtacos_retunegan_b200 as sb
    P = sb.preprocess
    good = (1e-4 * O.synth_noise(20000, 1)).astype(np.float32)
    good[2000:18000] += O.synth_speechlike(16000, 5)
    items = [("a", ("ka2 er2", "01"), good), ("b", ("pu3 pei2", "01"), np.zeros(300, np.float32) + 1e-3),
             ("c", ("wai4 sun1", "01"), np.ones(1, np.float32)), ("d", ("wan2 hua2", "01"), good[::-1].copy())]
    with pytest.warns(UserWarning, match="too short"):
        out = P.make_metadata_batch(items, str(tmp_path))
    assert out[1] is None and out[2] is None and out[0] is not None and out[3] is not None
    assert out[0][0] == "a" and os.path.exists(tmp_path / "mag-a.npy") and not os.path.exists(tmp_path / "mag-b.npy")
    with pytest.raises(ValueError):
        sb.core.yin(np.zeros(100, np.float32), 22050, 73.0, 590.0, 1024, 256)      # shorter than frame_length / 2 + 1
