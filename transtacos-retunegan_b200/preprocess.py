"""GPU-sharded ``make preprocess``: the reference's corpus preprocessing driver and its on-disk formats.

Mirrors ``transtacos/preprocess.py:16-41`` (``write_metadata``) and ``transtacos/datasets/databaker.py:25-124``
(``preprocess`` / ``make_metadata`` / ``parse_label_file``) with the per-utterance numpy work replaced by batched launches:

    reference (one process-pool task per utterance)      here (one rank per GPU, utterances in ragged batches)
    A.load_wav -> A.trim_silence -> A.align_wav          scipy wav read -> ONE frame_stats launch per batch -> host slices
    A.get_specs(y[:-1])  (mag, mel)                      ONE fused STFT+mel launch per batch (core.stft_features, ragged)
    A.get_c0(y[:-1])                                     ONE frame_stats launch per batch
    A.get_f0(y[:-1])     (librosa.yin)                   ONE yin launch per batch (or a caller-supplied ``f0_fn``)
    np.save mel-/mag-/f0-/c0-{name}.npy                  same files, same shapes / dtypes / memory order (see ``save_features``)

Files written (consumer contract: ``transtacos/data.py:153-161`` loads ``mel-{id}.npy`` / ``mag-{id}.npy`` and transposes them,
``retunegan/data.py:27-31,63-65`` reads ``wav_path.txt`` and the file lists):

    mel-{name}.npy  [n_mel, T]  float64, Fortran order   (``_normalize(...)`` of an F-ordered librosa.stft result)
    mag-{name}.npy  [n_freq, T] float64, Fortran order
    c0-{name}.npy   [T] float32;  f0-{name}.npy [T] float32
    train.txt / test.txt  ``name|prds|text`` lines;  stats.txt ``key<TAB>value``;  wav_path.txt

The features are computed in float32 on the GPU (1e-4 relative to the reference's float64, tests/test_gpu_parity.py) and
widened on the host, so a dataset written here is read by the reference's ``data.py`` unchanged.  The frame-major device
buffers ARE the Fortran-ordered ``[F, T]`` arrays, so the widening is the only host pass.  Utterances shard across ranks by
length (``sharding.shard_utterances``) with no data-path collective; rank 0 gathers the metadata tuples.
"""
from __future__ import annotations

import os
import warnings
import random
from collections import defaultdict
from re import compile as Regex
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import core, sharding
from . import transtacos_audio as A

DROPOUT_2SIGMA = True                                                   # datasets/databaker.py:18
PUNCT_KANJI_REGEX = Regex(r'，|。|、|：|；|？|！|（|）|“|”|…|—')           # datasets/databaker.py:22


def load_wav(path: str) -> np.ndarray:
    """``A.load_wav`` (transtacos/audio.py:31-33) for files that already are at ``hp.sample_rate``: float32 in (-1, 1), mono.
    librosa's ``kaiser_best`` resampler is not part of this path: another rate raises."""
    from scipy.io import wavfile
    sr, y = wavfile.read(path)
    if sr != A.hp.sample_rate:
        raise ValueError(f"{path}: sample rate {sr} != {A.hp.sample_rate}; resample the corpus first")
    if y.dtype.kind == "i":
        y = y.astype(np.float32) / float(np.iinfo(y.dtype).max + 1)   # soundfile / audioread scaling
    elif y.dtype.kind == "u":
        y = (y.astype(np.float32) - 128.0) / 128.0
    y = y.astype(np.float32)
    return y.mean(axis=1) if y.ndim == 2 else y


def save_features(out_dp: str, name: str, mag_fm: torch.Tensor, mel_fm: torch.Tensor, c0: torch.Tensor, f0=None,
                  dtype=np.float64) -> dict:
    """Write one utterance's files.  ``mag_fm [T, F]`` / ``mel_fm [T, M]`` are frame-major host tensors: their ``.T`` is the
    F-ordered ``[F, T]`` array the reference saves (np.save records ``fortran_order: True`` and writes the buffer as is)."""
    mag = np.asarray(mag_fm.numpy(), dtype=dtype).T
    mel = np.asarray(mel_fm.numpy(), dtype=dtype).T
    c0 = np.asarray(c0.numpy(), dtype=np.float32)
    np.save(os.path.join(out_dp, f'mel-{name}.npy'), mel, allow_pickle=False)
    np.save(os.path.join(out_dp, f'mag-{name}.npy'), mag, allow_pickle=False)
    np.save(os.path.join(out_dp, f'c0-{name}.npy'), c0, allow_pickle=False)
    stats = {'max_mel': mel.max(), 'min_mel': mel.min(), 'max_mag': mag.max(), 'min_mag': mag.min(),
             'max_c0': c0.max(), 'min_c0': c0.min()}
    if f0 is not None:
        f0 = np.asarray(f0, dtype=np.float32)
        np.save(os.path.join(out_dp, f'f0-{name}.npy'), f0, allow_pickle=False)
        stats.update({'max_f0': f0.max(), 'min_f0': f0.min()})
    return stats


def make_metadata_batch(items: Sequence[Tuple[str, Tuple[str, str], np.ndarray]], out_dp: str,
                        f0_fn: Optional[Callable[[np.ndarray], np.ndarray]] = None, dtype=np.float64) -> List[Optional[tuple]]:
    """``make_metadata`` (datasets/databaker.py:91-122) for a batch: ``items`` = (name, (text, prds), wav float32).
    Returns the reference's tuples ``(name, prds, text, len_text, len_wav, len_spec, stats)`` (None where the reference skips)."""
    hp = A.hp
    keep, ys = [], []
    for i, (name, (text, prds), y) in enumerate(items):
        if y is None or len(text.split(' ')) != len(prds):
            continue
        keep.append(i)
        ys.append(np.ascontiguousarray(y, np.float32))
    out: List[Optional[tuple]] = [None] * len(items)
    if not keep:
        return out
    plan = core.get_plan(hp)
    min_len = plan.n_fft // 4 + 2                        # reflect padding of the centred window needs n_fft/4 + 1 samples of y[:-1]
    ok = [j for j, y in enumerate(ys) if len(y) >= 2]
    trimmed = dict(zip(ok, A.trim_silence([ys[j] for j in ok]))) if ok else {}   # one launch for the batch
    good = []
    for j in range(len(ys)):
        y = A.align_wav(trimmed[j]) if j in trimmed else None
        if y is None or len(y) < min_len:
            # the reference handles clips one by one (librosa reflect-pads anything >= 2 samples, with a warning); a clip
            # shorter than the window support carries no usable frame: skip it instead of failing the whole batch
            warnings.warn(f"preprocess: utterance {items[keep[j]][0]!r} is too short after trimming "
                          f"({0 if y is None else len(y)} samples); skipped")
            continue
        ys[j] = y
        good.append(j)
    keep = [keep[j] for j in good]
    ys = [ys[j] for j in good]
    if not keep:
        return out
    cuts = [y[:-1] for y in ys]                          # datasets/databaker.py:102
    batch = core.SignalBatch(plan, cuts)
    sc = A.db_norm_scale(hp)
    mag, mel, _ = core.stft_features(plan, batch, preemph=hp.preemphasis, mag_scale=sc, mel_scale=sc)
    c0, _, _ = core.frame_stats(cuts, hp.win_length, hp.hop_length, want_zcr=False)
    f0_h = None
    if f0_fn is None:
        from .config import note_to_hz
        f0_h = core.yin(cuts, hp.sample_rate, note_to_hz(hp.rf0min), note_to_hz(hp.rf0max), hp.win_length, hp.hop_length)[0].cpu()
    mag_h, mel_h, c0_h = mag.cpu(), mel.cpu(), c0.cpu()   # one copy each for the batch
    o = 0
    for j, i in enumerate(keep):
        name, (text, prds), _ = items[i]
        T = int(batch.frames[j])
        len_wav = len(ys[j])
        assert len_wav == T * hp.hop_length               # datasets/databaker.py:108
        f0 = f0_fn(cuts[j]) if f0_fn is not None else f0_h[o:o + T].numpy()
        stats = save_features(out_dp, name, mag_h[o:o + T], mel_h[o:o + T], c0_h[o:o + T], f0, dtype)
        o += T
        out[i] = (name, prds, text, len(text.split(' ')), len_wav, T, stats)
    return out


def filter_and_aggregate(metadata: List[tuple], sample_rate: int):
    """2-sigma length filter and corpus statistics (datasets/databaker.py:39-88)."""
    metadata = [mt for mt in metadata if mt is not None]
    if DROPOUT_2SIGMA and metadata:
        tlens = np.asarray([mt[-4] for mt in metadata])
        alens = np.asarray([mt[-2] for mt in metadata])
        tL, tR = tlens.mean() - 2 * tlens.std(), tlens.mean() + 2 * tlens.std()
        aL, aR = alens.mean() - 2 * alens.std(), alens.mean() + 2 * alens.std()
        metadata = [mt for mt in metadata if tL <= mt[-4] <= tR and aL <= mt[-2] <= aR]
    len_text = np.asarray([mt[-4] for mt in metadata])
    len_wav = np.asarray([mt[-3] for mt in metadata])
    len_spec = np.asarray([mt[-2] for mt in metadata])
    agg = defaultdict(list)
    for mt in metadata:
        for k, v in mt[-1].items():
            agg[k].append(v)
    stats = {
        'total_examples': len(metadata),
        'total_hours': len_wav.sum() / sample_rate / (60 * 60),
        'min_len_txt': len_text.min(), 'max_len_txt': len_text.max(), 'avg_len_txt': len_text.mean(),
        'min_len_wav': len_wav.min(), 'max_len_wav': len_wav.max(), 'avg_len_wav': len_wav.mean(),
        'min_len_spec': len_spec.min(), 'max_len_spec': len_spec.max(), 'avg_len_spec': len_spec.mean(),
    }
    for k, v in agg.items():
        stats[k] = getattr(np.asarray(v), k[:k.find('_')])()   # 'max_mel' -> .max(), 'min_c0' -> .min()
    return [mt[:3] for mt in metadata], stats


def _prosody_marks(kanji: str) -> str:
    """One digit per character: '0' inside a word, the ``#n`` level that follows a character replaces its '0'."""
    marks: List[str] = []
    for ch in kanji.replace('#', ''):
        if not ch.isdigit():
            marks.append('0')
        elif marks:
            marks[-1] = ch
        else:
            marks.append(ch)
    return ''.join(marks)


def parse_label_file(fp: str) -> Dict[str, Tuple[str, str]]:
    """DataBaker ``000001-010000.txt`` -> {name: (pinyin, prosody digits)}; same result as datasets/databaker.py:125-160.
    The file alternates ``name<TAB>text with #n prosody marks`` and ``<TAB>pinyin`` lines; reading stops at the first blank
    header line, punctuation carries no prosody slot."""
    with open(fp, encoding='utf-8') as fh:
        lines = [ln.strip() for ln in fh.read().split('\n')]
    table: Dict[str, Tuple[str, str]] = {}
    for head, pinyin in zip(lines[0::2], lines[1::2] + ['']):
        if not head:
            break
        name, kanji = head.split('\t')
        table[name] = (pinyin.lower(), _prosody_marks(PUNCT_KANJI_REGEX.sub('', kanji)))
    return table


def write_metadata(metadata: List[tuple], stats: dict, wav_path: str, base_dir: str, out_dir: str = 'preprocessed',
                   shuffle: bool = True, split_ratio: float = 0.05, seed: Optional[int] = None) -> None:
    """``transtacos/preprocess.py:16-41``: train.txt / test.txt / stats.txt / wav_path.txt."""
    if shuffle:
        random.Random(A.hp.randseed if seed is None else seed).shuffle(metadata)
    out_path = os.path.join(base_dir, out_dir)
    os.makedirs(out_path, exist_ok=True)
    cp = int(len(metadata) * split_ratio)
    for fn, part in (('train.txt', metadata[cp:]), ('test.txt', metadata[:cp])):
        with open(os.path.join(out_path, fn), 'w', encoding='utf-8') as fh:
            for mt in part:
                fh.write('|'.join(str(x) for x in mt))
                fh.write('\n')
    with open(os.path.join(out_path, 'stats.txt'), 'w', encoding='utf-8') as fh:
        for k, v in stats.items():
            fh.write(f'{k}\t{v}\n')
    with open(os.path.join(out_path, 'wav_path.txt'), 'w', encoding='utf-8') as fh:
        fh.write(wav_path)


def preprocess_corpus(label_dict: Dict[str, Tuple[str, str]], wav_dp: str, out_dp: str, batch: int = 128,
                      f0_fn: Optional[Callable] = None, dtype=np.float64):
    """``datasets.databaker.preprocess`` over an explicit label dictionary.  Every rank processes its length-balanced shard
    (file size as the length proxy) in batches of ``batch`` utterances; rank 0 returns ``(metadata, stats)``, others (None, None)."""
    os.makedirs(out_dp, exist_ok=True)
    rank, world = sharding.rank_world()
    names = sorted(label_dict)
    sizes = [os.path.getsize(os.path.join(wav_dp, f'{n}.wav')) if os.path.exists(os.path.join(wav_dp, f'{n}.wav')) else 0
             for n in names]
    mine = sorted(sharding.shard_utterances(sizes, world)[rank])
    local: List[Optional[tuple]] = []
    for b0 in range(0, len(mine), batch):
        items = []
        for i in mine[b0:b0 + batch]:
            fp = os.path.join(wav_dp, f'{names[i]}.wav')
            items.append((names[i], label_dict[names[i]], load_wav(fp) if sizes[i] else None))
        local += make_metadata_batch(items, out_dp, f0_fn, dtype)
    if world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, local)
        if rank != 0:
            return None, None
        local = [mt for part in gathered for mt in part]
    order = {n: i for i, n in enumerate(names)}
    local = sorted((mt for mt in local if mt is not None), key=lambda mt: order[mt[0]])   # the reference's submission order
    return filter_and_aggregate(local, A.hp.sample_rate)


def preprocess(args, f0_fn: Optional[Callable] = None):
    """Drop-in for ``datasets.databaker.preprocess(args)`` (``args.base_dir``, ``args.out_dir``): returns (metadata, stats, wav_dp)."""
    wav_dp = os.path.join(args.base_dir, 'DataBaker', 'Wave')
    out_dp = os.path.join(args.base_dir, args.out_dir)
    label_dict = parse_label_file(os.path.join(args.base_dir, 'DataBaker', 'ProsodyLabeling', '000001-010000.txt'))
    metadata, stats = preprocess_corpus(label_dict, wav_dp, out_dp, f0_fn=f0_fn)
    return metadata, stats, wav_dp
