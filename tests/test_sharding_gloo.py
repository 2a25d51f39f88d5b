"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: image sharding with no data-path collective,
counter agreement, max-over-ranks timing, and bench.py's rank-0-only reference arm."""
import json
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from weaklysuperviseddl_b200 import sharding  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert sharding.rank_world() == (rank, world)
        n = 37
        mine = list(sharding.shard_indices(n, rank, world))
        # every rank processes its own images only: a fake "mask generation" whose result depends on the image index
        gen = torch.Generator().manual_seed(0)
        data = torch.rand(n, 8, 8, generator=gen)
        masks = {i: (data[i] >= 0.5).to(torch.uint8) for i in mine}
        near = sum(int(((data[i] - 0.5).abs() < 1e-2).sum()) for i in mine)
        total_near, total_masks = sharding.all_reduce_counters([near, len(masks)])
        ref_near = int(((data - 0.5).abs() < 1e-2).sum())
        assert total_masks == n and total_near == ref_near
        # the union of the shards is a partition of the image set
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        flat = sorted(i for part in gathered for i in part)
        assert flat == list(range(n))
        assert all(i % world == r for r, part in enumerate(gathered) for i in part)
        # timing rule: the slowest rank
        assert sharding.max_over_ranks(1.0 + rank) == float(world)
        # batches of a loader
        got = [i for i, _ in sharding.shard_batches(range(10), rank, world)]
        assert got == list(range(rank, 10, world))
        torch.save({i: m for i, m in masks.items()}, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_image_sharding_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    merged = {}
    for r in range(world):
        merged.update(torch.load(os.path.join(tmp_path, f"rank{r}.pt")))
    gen = torch.Generator().manual_seed(0)
    data = torch.rand(37, 8, 8, generator=gen)
    assert sorted(merged) == list(range(37))
    for i, m in merged.items():  # sharded result == single-process result, image by image
        assert torch.equal(m, (data[i] >= 0.5).to(torch.uint8))


def _hooks_for(indices):
    """Synthetic hooks, seed = image index, tiny shapes."""
    acts, grads = [], []
    for (C, h, w) in ((6, 5, 5), (4, 3, 3)):
        a, g = [], []
        for i in indices:
            gen = torch.Generator().manual_seed(1000 + i)
            a.append(torch.randn(C, h, w, generator=gen))
            g.append(torch.randn(C, h, w, generator=gen) * 1e-3)
        acts.append(torch.stack(a))
        grads.append(torch.stack(g))
    return acts, grads


def _oracle_mask_fn(acts, grads):
    from oracle import wsdl_oracle as O

    cam = O.layercam_from_hooks(acts, grads, (24, 20))
    return ((cam >= 0.3) & (cam > 0)).to(torch.uint8), ((cam - 0.3).abs() < 1e-6).sum().reshape(1)


def _sharded_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from weaklysuperviseddl_b200.PsuedoMasks import generate_pseudo_masks_sharded

        res = generate_pseudo_masks_sharded(_hooks_for, 23, (24, 20), chunk=4, mask_fn=_oracle_mask_fn, count_foreground=True)  # rank/world from the group
        torch.save(res, os.path.join(out_dir, f"sharded{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_sharded_pseudo_masks_union_equals_single_rank(tmp_path):
    """generate_pseudo_masks_sharded (the config-3 product entry) at world size 2 over gloo: the union of the two shards
    is the single-rank result bit for bit, image by image, and both ranks agree on the all-reduced counters.  The fused
    launch is replaced by the oracle (mask_fn) so that the host logic runs without a GPU; the GPU twin of this test is
    tests/test_gpu_configs.py::test_sharded_pseudo_masks_on_device."""
    from weaklysuperviseddl_b200.PsuedoMasks import generate_pseudo_masks_sharded

    world = 2
    mp.spawn(_sharded_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    single = generate_pseudo_masks_sharded(_hooks_for, 23, (24, 20), chunk=5, rank=0, world=1, mask_fn=_oracle_mask_fn, count_foreground=True)
    assert single["indices"].tolist() == list(range(23)) and single["counters"]["masks"] == 23
    seen = {}
    for r in range(world):
        part = torch.load(os.path.join(tmp_path, f"sharded{r}.pt"))
        assert part["counters"] == single["counters"]  # summed over ranks == the single-rank totals
        assert part["indices"].tolist() == list(range(r, 23, world))
        for j, i in enumerate(part["indices"].tolist()):
            seen[i] = part["masks"][j]
    assert sorted(seen) == list(range(23))
    for i in range(23):
        assert torch.equal(seen[i], single["masks"][i]), i
    # a sink takes the masks instead of the return value
    got = {}
    res = generate_pseudo_masks_sharded(_hooks_for, 7, (24, 20), chunk=3, rank=0, world=1, mask_fn=_oracle_mask_fn,
                                        sink=lambda idx, m: got.update({i: m[j] for j, i in enumerate(idx)}))
    assert res["masks"] is None and sorted(got) == list(range(7)) and torch.equal(got[3], single["masks"][3])


def test_chunking_and_argument_checks():
    idx = list(sharding.shard_indices(10, 1, 4))
    assert idx == [1, 5, 9]
    assert sharding.chunk(list(range(7)), 3) == [[0, 1, 2], [3, 4, 5], [6]]
    import pytest

    with pytest.raises(ValueError):
        sharding.shard_indices(4, 2, 2)
    with pytest.raises(ValueError):
        sharding.chunk([1], 0)
    assert sharding.all_reduce_counters([3, 4]) == [3, 4]  # no process group: identity


def test_bench_reference_arm_rank0_only():
    """`bench.py --impl reference` under a 2-rank launch: rank 0 prints the line, rank 1 exits 0 silently."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    outs = []
    for rank in (0, 1):
        e = dict(env, RANK=str(rank), LOCAL_RANK=str(rank))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                            "--steps", "1", "--warmup", "0"], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout.strip())
    assert outs[1] == ""
    line = json.loads(outs[0].splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
