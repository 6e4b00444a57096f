"""CPU, world_size 2 over gloo: the data-parallel rule of the trainer — identical draws on
every rank, each rank keeps its batch slice, gradients are all-reduced and scaled by 1/R — reproduces
the single-process gradient of the global batch.  Uses the engine with the torch test double."""
import contextlib
import io
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from leco_b200.lora import LoRANetwork
from leco_b200.unet import SPECS, EngineUNet
from oracle.unet_ref import build_unet
from tests import torch_backend


def _grads(rank, world, bg=2):
    eng = EngineUNet(SPECS["tiny21"], backend=torch_backend)
    eng.load_state_dict(build_unet("tiny21").state_dict())
    eng._act_dtype = torch.float32
    eng.requires_grad_(False)
    torch.manual_seed(1234)                       # identical on every rank
    with contextlib.redirect_stdout(io.StringIO()):
        net = LoRANetwork(eng, rank=4, alpha=1.0)
    g = torch.Generator().manual_seed(5)
    for l in net.unet_loras:
        l.lora_up.weight.data = 0.05 * torch.randn(l.lora_up.weight.shape, generator=g)
    noise = torch.randn((bg, 4, 8, 8), generator=g)          # GLOBAL batch drawn identically, then sliced
    goal = torch.randn((bg, 4, 8, 8), generator=g)
    ctx = torch.randn((1, 77, 128), generator=g)
    bl = bg // world
    sl = slice(rank * bl, (rank + 1) * bl)
    with net:
        y = eng(noise[sl], torch.tensor(500), encoder_hidden_states=ctx.expand(bl, -1, -1)).sample
    loss = torch.nn.functional.mse_loss(y, goal[sl])         # mean over the LOCAL batch
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for l in net.unet_loras for p in l.parameters()])
    return flat, loss.detach()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    flat, loss = _grads(rank, world)
    dist.all_reduce(flat)
    dist.all_reduce(loss)
    flat /= world
    loss /= world
    if rank == 0:
        torch.save({"flat": flat, "loss": loss}, out)
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_process(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.set_num_threads(4)
    want_flat, want_loss = _grads(0, 1)
    assert abs(got["loss"].item() - want_loss.item()) < 1e-6 * abs(want_loss.item()) + 1e-9
    rel = (got["flat"] - want_flat).norm() / want_flat.norm()
    assert rel < 1e-4, rel


def _seed_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from leco_b200 import train_lora
    torch.manual_seed(1000 + rank)                 # ranks start out different, as un-seeded processes do
    seed = train_lora.share_seed(torch.device("cpu"))
    draws = (torch.randint(0, 1 << 30, (4,)), torch.randn(3))
    torch.save({"seed": seed, "draws": draws}, out + str(rank))
    dist.destroy_process_group()


def test_ranks_share_rank0_seed(tmp_path):
    """leco_b200.train_lora.train() under torchrun: every rank installs rank 0's seed before any draw."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "seed")
    mp.spawn(_seed_worker, args=(2, port, out), nprocs=2, join=True)
    a, b = torch.load(out + "0"), torch.load(out + "1")
    assert a["seed"] == b["seed"] == 1000
    assert torch.equal(a["draws"][0], b["draws"][0]) and torch.equal(a["draws"][1], b["draws"][1])
    from leco_b200 import train_lora
    assert train_lora.dist_env({"RANK": "3", "WORLD_SIZE": "8", "LOCAL_RANK": "3"}) == (3, 8, 3)
    assert train_lora.dist_env({}) == (0, 1, 0)
