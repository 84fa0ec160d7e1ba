"""N > 1 host logic on the CPU: world_size-2 gloo run of the sharded stage loop over the TEST-ONLY CPU
simulator, compared with the single-process run.  The sharding rule and the atlas layout are the
product's (include/dpe_b200.h, "several GPUs"): reference views split into contiguous balanced blocks
(dpe_shard_range, called through the C ABI here), atlas slot of a view = local index * n_ranks + owner
rank, and one in-place all-gather per view slot — what dpe_stage_begin queues on NCCL."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT / "dpe-mvs_b200", ROOT / "oracle", ROOT / "tests"):
        sys.path.insert(0, str(p))
    import hostsim
    from scenes import small_scene
    import simpipe
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec, grays, cams, drs, pairs, gt = small_scene("c1", 0.125, None)      # 80 x 60, 5 views
    V = len(grays)
    H, W = grays[0].shape
    ns = 2
    sizes = simpipe.level_sizes(W, H, ns)
    pyr = [[hostsim.resize_linear(g.astype(np.float32), *sizes[k]) if sizes[k] != (W, H) else g.astype(np.float32) for k in range(ns)] for g in grays]
    import capi
    spr = (V + world - 1) // world
    first, count = capi.shard_range(V, world, rank)
    blocks = [capi.shard_range(V, world, r) for r in range(world)]

    def slot_of(v):
        r = next(r for r, (f, c) in enumerate(blocks) if f <= v < f + c)
        return (v - blocks[r][0]) * world + r

    state = {}
    atlas_front = {k: torch.zeros(world * spr, sizes[k][1], sizes[k][0]) for k in range(ns)}
    for si, (k, p) in enumerate(hostsim.stage_schedule(ns)[:4]):
        back = torch.zeros_like(atlas_front[k])
        for v in range(first, first + count):
            ids = [v] + list(pairs[v])
            sd = [atlas_front[k][slot_of(i)].numpy() for i in pairs[v]] if p.geom_consistency else None
            prev = None if v not in state else (state[v]["planes"], state[v]["state"], state[v]["selected"])
            r = hostsim.run_stage([pyr[i][k] for i in ids], [cams[i] for i in ids], drs[v], (W, H), p, 7, view=v,
                                  stage_counter=si, prev=prev, src_depths=sd)
            state[v] = r
            back[slot_of(v)] = torch.from_numpy(r["depth"])
        # one in-place all-gather per view slot i: ranks contribute slot i * world + rank
        for i in range(spr):
            chunks = [torch.empty(1, *back.shape[1:]) for _ in range(world)]
            dist.all_gather(chunks, back[i * world + rank:i * world + rank + 1].contiguous())
            back[i * world:(i + 1) * world] = torch.cat(chunks, 0)
        atlas_front[k] = back
    # report in view order
    out = torch.stack([atlas_front[0][slot_of(v)] for v in range(V)])
    pad = torch.stack([atlas_front[0][s] for s in range(world * spr) if s not in {slot_of(v) for v in range(V)}]) if world * spr > V else torch.zeros(0, *out.shape[1:])
    np.save(Path(out_dir) / f"pad_rank{rank}.npy", pad.numpy())
    np.save(Path(out_dir) / f"atlas_rank{rank}.npy", out.numpy())
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_sharding_matches_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a0 = np.load(tmp_path / "atlas_rank0.npy")
    a1 = np.load(tmp_path / "atlas_rank1.npy")
    assert np.array_equal(a0, a1), "every rank must hold the same gathered atlas"
    # single process, same schedule: identical depth maps (results do not depend on the sharding)
    port2 = port + 1
    single = tmp_path / "single"
    single.mkdir()
    mp.spawn(_worker, args=(1, port2, str(single)), nprocs=1, join=True)
    s = np.load(single / "atlas_rank0.npy")
    assert np.array_equal(s, a0)
    assert (np.load(tmp_path / "pad_rank0.npy") == 0).all()        # padded slots stay empty
