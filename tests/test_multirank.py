"""World-size-2 gloo test (CPU) of the sweep sharding + final gather used by bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_points, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import warnings
    warnings.filterwarnings('ignore')
    import circuits as cs
    from emul_kernels import EmulKernels
    from sqcircuit_b200.sharding import gather_to_root, shard
    cr, lp, _, _ = cs.zero_pi(EmulKernels())
    cr.set_trunc_nums([9, 5])
    fluxes = np.linspace(0.0, 0.5, n_points)
    mine = shard(fluxes, rank, world)
    res = cr.sweep_diag(3, loop_fluxes={lp: mine})
    full = gather_to_root(res.efreqs.contiguous(), n_points)
    if rank == 0:
        torch.save(full, out)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sweep_equals_single_rank(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import circuits as cs
    from emul_kernels import EmulKernels
    n_points = 7      # odd: shards of 4 and 3 points
    out = str(tmp_path / 'gathered.pt')
    mp.spawn(_worker, args=(2, _free_port(), n_points, out), nprocs=2, join=True)
    got = torch.load(out)
    cr, lp, _, _ = cs.zero_pi(EmulKernels())
    cr.set_trunc_nums([9, 5])
    ref = cr.sweep_diag(3, loop_fluxes={lp: np.linspace(0.0, 0.5, n_points)}).efreqs
    assert got.shape == (n_points, 3)
    assert torch.allclose(got, ref, rtol=1e-10, atol=0)


def test_shard_bounds_cover_everything():
    from sqcircuit_b200.sharding import shard_bounds
    for n in (1, 7, 4096, 4099):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
