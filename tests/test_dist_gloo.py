"""World-size-2 CPU test (gloo) of the multi-GPU plumbing: utterances shard across ranks with no data-path
collective; only the corpus min/max and the scalar loss mean are reduced."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import transtacos_retunegan_b200 as sb
    from oracle import spectral_oracle as O
    lens = [256 * t - 1 for t in (12, 30, 9, 21, 17, 26, 14)]
    mine = sb.sharding.shard_utterances(lens, world)[rank]
    # stand-in for the per-rank GPU work: the checker computes this rank's features on the CPU
    stats, sums = [], {}
    for i in mine:
        S, M = O.tt_get_specs(O.synth_noise(lens[i], 100 + i))
        stats.append((S.min(), S.max()))
        sums[i] = float(S.sum() + M.sum())
    lo, hi = sb.sharding.reduce_stats(min(s[0] for s in stats), max(s[1] for s in stats))
    local_loss = torch.tensor(float(rank + 1))
    mean = sb.sharding.reduce_mean_scalar(local_loss)
    assert sb.sharding.rank_world() == (rank, world)
    # the in-kernel exchange over NVLink peer memory is an NCCL-box path: on gloo / without CUDA ddp_reduce takes the all-reduce
    assert sb.loss.ddp_loss_reducer() is None and sb.loss.ddp_reduce_path() == "NCCL all-reduce of one scalar"
    red = sb.loss._GlobalMean.apply(local_loss.clone().requires_grad_(True))     # value = mean over the ranks, gradient = local
    assert abs(float(red) - (world + 1) / 2) < 1e-12
    q.put((rank, mine, sums, lo, hi, float(mean)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    all_idx = sorted(i for r in res for i in r[1])
    assert all_idx == list(range(7))                       # every utterance on exactly one rank
    # single-process answer
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import spectral_oracle as O
    lens = [256 * t - 1 for t in (12, 30, 9, 21, 17, 26, 14)]
    lo = min(O.tt_get_specs(O.synth_noise(lens[i], 100 + i))[0].min() for i in range(7))
    hi = max(O.tt_get_specs(O.synth_noise(lens[i], 100 + i))[0].max() for i in range(7))
    for r in res:
        assert abs(r[3] - lo) < 1e-12 and abs(r[4] - hi) < 1e-12      # reduced stats identical on both ranks
        assert abs(r[5] - 1.5) < 1e-12                                  # mean of rank losses (1, 2)
    merged = {}
    for r in res:
        merged.update(r[2])
    for i in range(7):
        S, M = O.tt_get_specs(O.synth_noise(lens[i], 100 + i))
        assert merged[i] == float(S.sum() + M.sum())                    # sharding never changes a result


def _pp_worker(rank, world, port, base, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import transtacos_retunegan_b200 as sb
    P = sb.preprocess
    seen = []

    def fake_batch(items, out_dp, f0_fn=None, dtype=None):   # stand-in for the per-rank GPU work (no GPU here)
        out = []
        for name, (text, prds), y in items:
            seen.append(name)
            st = {'max_mel': float(len(y)), 'min_mel': -1.0, 'max_mag': 1.0, 'min_mag': -5.6, 'max_c0': .5, 'min_c0': 0.}
            out.append((name, prds, text, len(prds), len(y), len(y) // 256, st))
        return out

    P.make_metadata_batch = fake_batch
    P.DROPOUT_2SIGMA = False
    labels = {f"{i:06d}": ("a1 b2 c3", "004") for i in range(1, 10)}
    meta, stats = P.preprocess_corpus(labels, os.path.join(base, "Wave"), os.path.join(base, "out"), batch=2)
    q.put((rank, seen, meta, None if stats is None else {k: float(v) for k, v in stats.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_preprocess_driver(tmp_path):
    """preprocess_corpus: every wav on exactly one rank, rank 0 gathers the tuples in the reference's order."""
    from scipy.io import wavfile
    (tmp_path / "Wave").mkdir()
    lens = {}
    for i in range(1, 10):
        L = 256 * (10 + 7 * i % 23)
        lens[f"{i:06d}"] = L
        wavfile.write(tmp_path / "Wave" / f"{i:06d}.wav", 22050, np.zeros(L, np.float32))
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_pp_worker, args=(r, world, port, str(tmp_path), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res[0][1] + res[1][1]) == sorted(lens) and res[0][1] and res[1][1]
    assert abs(sum(lens[n] for n in res[0][1]) - sum(lens[n] for n in res[1][1])) <= max(lens.values())   # length-balanced
    assert res[1][2] is None and res[1][3] is None
    assert [m[0] for m in res[0][2]] == sorted(lens) and res[0][2][0] == ("000001", "004", "a1 b2 c3")
    assert res[0][3]['total_examples'] == 9 and res[0][3]['max_mel'] == float(max(lens.values()))
