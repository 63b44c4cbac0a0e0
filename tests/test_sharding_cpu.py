"""CPU tier: the multi-GPU host logic (shot sharding, gather of predictions, sum of mistake counts) under gloo with
world_size 2 and 3. The per-rank "decoder" here is the CPU oracle -- this test is about the plumbing, not the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dem_util import read_dem, NumpyDemSampler

import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))


def test_shard_bounds_cover_everything():
    from dist_sharding import shard_bounds
    for n in [0, 1, 7, 64, 1000003]:
        for w in [1, 2, 3, 8]:
            pieces = [shard_bounds(n, r, w) for r in range(w)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(pieces[k][1] == pieces[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in pieces]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, name, shots, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, 'tests'))
    sys.path.insert(0, os.path.join(root, 'tools'))
    import oracle
    from dist_sharding import decode_sharded
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    text = read_dem(name)
    dets, obs = NumpyDemSampler(text).sample(shots, seed=21)
    dec = oracle.OracleDecoder(text)
    full, mistakes = decode_sharded(dec.predict_obs_flips_from_dets_bit_packed, dets, actual_obs=obs, dist=dist)
    np.save(os.path.join(out_dir, 'pred_%d.npy' % rank), full)
    np.save(os.path.join(out_dir, 'mist_%d.npy' % rank), np.array([mistakes]))
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_sharded_decode_matches_single_process(built, tmp_path, world):
    import oracle
    name, shots = 'midout_color_code_d5_r10_p1000', 1001
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, name, shots, str(tmp_path)), nprocs=world, join=True)
    text = read_dem(name)
    dets, obs = NumpyDemSampler(text).sample(shots, seed=21)
    want = oracle.OracleDecoder(text).predict_obs_flips_from_dets_bit_packed(dets)
    want_mistakes = int((want[:, 0].astype(np.uint64) != obs).sum())
    for r in range(world):
        got = np.load(os.path.join(str(tmp_path), 'pred_%d.npy' % r))
        assert got.shape == want.shape and (got == want).all()
        assert int(np.load(os.path.join(str(tmp_path), 'mist_%d.npy' % r))[0]) == want_mistakes
