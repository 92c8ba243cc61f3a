"""
Multi-GPU host logic: sample shares, frame partition, disjoint RNG
streams, and the histogram reduce -- exercised on CPU with two gloo ranks; the
NCCL path runs on the GPU box (bench.py --gpus N, test marked gpu below).
"""
import os
import socket

import numpy as np
import pytest


def test_sample_share_covers_frame_exactly():
    from cuburn_b200 import multigpu
    for total in (1, 65535, 65536, 65537, 1920 * 1080 * 2000, 3840 * 2160 * 4000 + 17):
        for world in (1, 2, 4, 8):
            parts = [multigpu.sample_share(total, r, world) for r in range(world)]
            assert sum(n for _, n in parts) == total
            pos = 0
            for first, n in parts:
                assert first % 65536 == 0
                if n:
                    assert first == pos
                    pos += n
            if total >= world * 65536:
                sizes = [n for _, n in parts]
                assert max(sizes) - min(sizes) <= 65536


def test_partition_frames_round_robin():
    from cuburn_b200 import multigpu
    frames = list(range(720))
    parts = [multigpu.partition_frames(frames, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == frames
    assert all(len(p) == 90 for p in parts) and parts[3][:3] == [3, 11, 19]


def test_rank_seeds_are_disjoint(built):
    from cuburn_b200 import multigpu, mwc
    world = 8
    seeds = [multigpu.make_rank_seeds(r, world, host_seed=5, nstreams=4096) for r in range(world)]
    mults = [set(s[:, 0].tolist()) for s in seeds]
    for a in range(world):
        for b in range(a + 1, world):
            assert not (mults[a] & mults[b])
    table = set(mwc.load_mults().tolist())
    assert all(m <= table for m in mults)
    assert not np.array_equal(seeds[0][:, 1], seeds[1][:, 1])
    again = multigpu.make_rank_seeds(3, world, host_seed=5, nstreams=4096)
    assert np.array_equal(again, seeds[3])


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from cuburn_b200 import multigpu, samples
    from oracle import flame_ref as R
    r, w, _ = multigpu.init_process_group('gloo')
    assert (r, w) == (rank, world)
    gnm = samples.g3()
    ev = R.GenomeEval(gnm, 160, 90, 0.5, 0.0)
    total = 160 * 90 * 40
    first, n = multigpu.sample_share(total, rank, world)
    seeds = multigpu.make_rank_seeds(rank, world, host_seed=9, nstreams=32768)
    pal, _ = R.palette_table(gnm, ev.ts, 0.0, multigpu.make_rank_seeds(0, world, 9, 32768))
    hist, _ = R.iterate(ev, pal, seeds[16384:], n, ntraj=64, nthreads=1)
    np.save(os.path.join(out_dir, 'hist_%d.npy' % rank), hist)
    summed = multigpu.reduce_host_hist(hist.copy(), root=0)
    if rank == 0:
        np.save(os.path.join(out_dir, 'reduced.npy'), summed)
        np.save(os.path.join(out_dir, 'share.npy'), np.array([first, n]))
    frames = multigpu.partition_frames(list(range(10)), rank, world)
    assert frames == list(range(rank, 10, world))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_hist_reduce_world2(built, tmp_path):
    """The reduce equals the sum of the per-rank histograms; shares add up."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    h0 = np.load(tmp_path / 'hist_0.npy')
    h1 = np.load(tmp_path / 'hist_1.npy')
    red = np.load(tmp_path / 'reduced.npy')
    assert np.array_equal(red, h0 + h1)
    assert not np.array_equal(h0, h1)            # independent streams
    total = 160 * 90 * 40
    inside = red[..., 3].sum()
    assert 0.9 * total < inside <= total


@pytest.mark.gpu
def test_hist_view_is_zero_copy(native, built):
    """torch sees the library's device buffer through __cuda_array_interface__."""
    import torch
    N = native
    buf = N.DeviceBuffer(4 * 1024)
    N.fill32(buf, 1024, np.float32(1.5))
    N.check(N.lib().cb_device_sync())
    t = torch.as_tensor(buf.view((1024,), '<f4'), device='cuda')
    assert t.data_ptr() == buf.ptr and float(t.sum()) == 1536.0
    t += 1
    torch.cuda.synchronize()
    assert np.all(N.from_device(buf, (1024,), np.float32) == 2.5)
