"""CPU test of the multi-GPU plumbing with world_size 2 over gloo: the slot partition, the seed
slicing and the single sum-reduce of the per-pixel accumulation image.  Each rank gets its slots'
accumulators from the oracle (slots are independent, so rendering only a rank's slots equals
taking them out of a full render), reduces, and rank 0 compares with the 1-process image."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS, ROWS, RPP = 24, 16, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib

    import torch
    import torch.distributed as dist

    import util
    from oracle import refcl as OR
    multi = importlib.import_module("2015-raytracing_b200").multi
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o_scene, _ = util.make_scene_pair(os.path.join(outdir, "r%d" % rank), COLS, ROWS, mesh_uv=(16, 8), mesh_nslabs=6)
        lib = OR.load_best()
        pixels, total = COLS * ROWS, COLS * ROWS * RPP
        seeds = OR.make_seeds(total, 5)
        st, pix_full, _ = OR.a10_render(lib, o_scene, COLS, ROWS, RPP, passes=1, seeds=seeds)
        begin, count = multi.slot_range(rank, world, RPP)
        mine = st.acu.reshape(pixels, RPP, 4)[:, begin:begin + count]
        part = np.zeros((pixels, 4), np.float32)
        for k in range(count):
            part += mine[:, k]
        t = torch.from_numpy(part.copy())
        multi.reduce_accum(t, dst=0)
        my_seeds = multi.local_seeds(st.seeds, pixels, RPP, begin, count)
        gathered = [None] * world
        dist.all_gather_object(gathered, (begin, count, my_seeds))
        if rank == 0:
            full = np.zeros((pixels, 4), np.float32)
            for k in range(RPP):
                full += st.acu.reshape(pixels, RPP, 4)[:, k]
            got = t.numpy()
            np.save(os.path.join(outdir, "err.npy"), np.array([np.abs(got - full).max(), np.abs(full).max()]))
            pix = multi.accum_to_pixel(got, RPP, 1)
            np.save(os.path.join(outdir, "pixdiff.npy"), np.array([np.abs(pix.astype(int) - pix_full.reshape(-1, 4).astype(int)).max()]))
            merged = multi.merge_seeds(gathered, pixels, RPP)
            np.save(os.path.join(outdir, "seeds_ok.npy"), np.array([int(np.array_equal(merged, st.seeds))]))
    finally:
        dist.destroy_process_group()


def test_slot_range_partitions_every_slot(rt):
    for rpp in (1, 4, 16, 100, 256):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                b, c = rt.multi.slot_range(r, world, rpp)
                seen += list(range(b, b + c))
            assert seen == list(range(rpp))
    with pytest.raises(ValueError):
        rt.multi.slot_range(2, 2, 4)


def test_seed_slicing_round_trip(rt):
    pixels, rpp = 7, 9
    g = np.arange(pixels * rpp, dtype=np.int32)
    parts = []
    for r in range(4):
        b, c = rt.multi.slot_range(r, 4, rpp)
        parts.append((b, c, rt.multi.local_seeds(g, pixels, rpp, b, c)))
    assert parts[1][2][:3].tolist() == [3, 4, 12]
    assert np.array_equal(rt.multi.merge_seeds(parts, pixels, rpp), g)


def test_accum_to_pixel_matches_oracle_copyToPixel(rt, oracle_lib):
    rng = np.random.Generator(np.random.PCG64(3))
    acu = rng.uniform(-0.2, 3.0, size=(50 * 4, 4)).astype(np.float32)
    want = np.zeros((50, 4), np.uint8)
    oracle_lib.a10_copyToPixel(want, acu, float(np.float32(1.0 / (4 * 3))), 50, 4)
    summed = np.zeros((50, 4), np.float32)
    for k in range(4):
        summed += acu.reshape(50, 4, 4)[:, k]
    assert np.array_equal(rt.multi.accum_to_pixel(summed, 4, 3), want)


@pytest.mark.timeout(300)
def test_two_rank_reduce_over_gloo():
    import torch.multiprocessing as mp
    outdir = tempfile.mkdtemp(prefix="rt_gloo_")
    mp.spawn(_worker, args=(2, _free_port(), outdir), nprocs=2, join=True)
    err, scale = np.load(os.path.join(outdir, "err.npy"))
    assert err <= 4e-6 * max(scale, 1.0), "reduced image differs from the 1-process sum beyond fp32 reassociation"
    assert np.load(os.path.join(outdir, "pixdiff.npy"))[0] <= 1
    assert np.load(os.path.join(outdir, "seeds_ok.npy"))[0] == 1
