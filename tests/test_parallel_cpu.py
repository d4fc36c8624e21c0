"""world_size-2 gloo tests (CPU) of the coordinate-sharded data-parallel step: sharded gradients
summed with one flat all-reduce equal the full-batch gradients, and replicas stay identical."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import siren_oracle as so


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from siren_mri_b200 import modules, parallel
    torch.manual_seed(0)
    m = modules.SingleBVPNet(in_features=2, out_features=1, hidden_features=64, num_hidden_layers=2).double()
    x = torch.from_numpy(so.make_coords(1, n, 2, seed=1)).double()
    gt = torch.from_numpy(so.make_coords(1, n, 1, seed=2)).double()
    xs, gts = parallel.shard_coords(x, gt)
    assert xs.shape[1] in (n // world, n // world + 1)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    for _ in range(3):
        opt.zero_grad()
        out = m({"coords": xs})
        loss = ((out["model_out"] - gts) ** 2).sum() / n        # GLOBAL normalisation on every rank
        loss.backward()
        parallel.allreduce_gradients(list(m.parameters()))
        opt.step()
    torch.save({k: v.clone() for k, v in m.state_dict().items()}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_sharded_step_equals_full_batch_step(tmp_path):
    n, world = 1001, 2                                     # ragged: shards of 501 and 500
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    from siren_mri_b200 import modules
    torch.manual_seed(0)
    m = modules.SingleBVPNet(in_features=2, out_features=1, hidden_features=64, num_hidden_layers=2).double()
    x = torch.from_numpy(so.make_coords(1, n, 2, seed=1)).double()
    gt = torch.from_numpy(so.make_coords(1, n, 1, seed=2)).double()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    for _ in range(3):
        opt.zero_grad()
        loss = ((m({"coords": x})["model_out"] - gt) ** 2).sum() / n
        loss.backward()
        opt.step()
    r0 = torch.load(os.path.join(tmp_path, "rank0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "rank1.pt"))
    for k, v in m.state_dict().items():
        assert torch.equal(r0[k], r1[k]), k                # replicas bit-identical, no broadcast needed
        assert torch.allclose(r0[k], v, rtol=1e-9, atol=1e-12), k


def test_shard_bounds_cover_everything():
    from siren_mri_b200.parallel import shard_bounds
    for n in (1, 7, 262144, 250000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _reducer_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from siren_mri_b200 import parallel
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    red = parallel.PeerGradientReducer(net.parameters(), average=True)
    assert not red.fused                                      # CPU / gloo: the flat buffer + one all_reduce
    x = torch.full((4, 5), float(rank + 1))
    for it in range(2):
        red.zero_grad()
        net(x).pow(2).sum().backward()
        assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in net.parameters())      # still views of the buffer
        red.reduce()
    torch.save([p.grad.clone() for p in net.parameters()], os.path.join(out_dir, "g%d.pt" % rank))
    dist.destroy_process_group()


def test_peer_gradient_reducer_host_logic(tmp_path):
    """PeerGradientReducer off the GPU (gloo, world 2): gradients accumulate into the flat buffer through the views,
    reduce() leaves the MEAN over the ranks on every rank (what DDP's reducer leaves), identical on both."""
    world = 2
    mp.spawn(_reducer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(os.path.join(tmp_path, "g0.pt"))
    g1 = torch.load(os.path.join(tmp_path, "g1.pt"))
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    want = [torch.zeros_like(p) for p in net.parameters()]
    for r in range(world):
        net.zero_grad()
        net(torch.full((4, 5), float(r + 1))).pow(2).sum().backward()
        for w, p in zip(want, net.parameters()):
            w += p.grad / world
    for a, b, w in zip(g0, g1, want):
        assert torch.equal(a, b)
        assert torch.allclose(a, w, rtol=1e-6, atol=1e-7)
