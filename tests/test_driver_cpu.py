"""CPU suite, part 3: the batch-sharded driver on a 2-rank gloo group (host logic only)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multimodal_vqvae_compression_audio_tactile_b200 import driver


def test_shard_bounds_cover_everything():
    for n in (0, 1, 5, 8, 21, 64):
        for w in (1, 2, 3, 8):
            spans = [driver.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_forward(a, t, books_use):
    y = a * 2 + t
    idx = (a.sum(dim=(1, 2)) * 1000).long().view(-1, 1, 1).expand(-1, 3, 5).contiguous()
    return y, idx


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    a = torch.rand(7, 1, 16, generator=g)
    t = torch.rand(7, 1, 16, generator=g)
    sc = driver.ShardedCodec(_fake_forward)
    y_local, idx_all, y_all = sc.run(a, t, gather_y=True)
    y_ref, idx_ref = _fake_forward(a, t, None)
    lo, hi = sc.local_slice(7)
    ok = torch.equal(y_local, y_ref[lo:hi]) and torch.equal(idx_all, idx_ref) and torch.equal(y_all, y_ref)
    # bench.py's use: the hot path on the rank's own shard (no communication), one gather after the last step
    y2, idx2 = sc.run_local(a[lo:hi], t[lo:hi])
    counts = [b - a_ for a_, b in (driver.shard_bounds(7, world, r) for r in range(world))]
    ok = ok and torch.equal(y2, y_ref[lo:hi]) and torch.equal(sc.gather_indices(idx2, counts), idx_ref)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_codec_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_bench_picks_the_traffic_capture_of_its_micro_batch(tmp_path):
    """bench.py's roofline.traffic must come from an ncu capture taken at the micro-batch it runs (bytes per launch scale
    with it), and only for the bf16x3 family the captures cover."""
    import json
    import bench
    (tmp_path / "r02_ncu_dram_traffic_conv_mb128.json").write_text(json.dumps({"x3": {"dram_bytes_per_launch": 2.0}}))
    (tmp_path / "r02_ncu_dram_traffic_conv_mb64.json").write_text(json.dumps({"x3": {"dram_bytes_per_launch": 1.0}}))
    assert bench.pick_traffic(128, "conv_tc_x3", str(tmp_path)) == (2.0, "profiles/r02_ncu_dram_traffic_conv_mb128.json")
    assert bench.pick_traffic(64, "conv_tc_x3", str(tmp_path)) == (1.0, "profiles/r02_ncu_dram_traffic_conv_mb64.json")
    assert bench.pick_traffic(32, "conv_tc_x3", str(tmp_path)) == (None, None)
    assert bench.pick_traffic(128, "conv_tc", str(tmp_path)) == (None, None)
    # the committed captures exist for the default step (two programs of 128) and for round 1's 64
    assert bench.pick_traffic(128, "conv_tc_x3")[0] and bench.pick_traffic(64, "conv_tc_x3")[0]
    assert bench.parse.__defaults__ is None           # defaults live in argparse: check them there
    import sys
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        a = bench.parse()
    finally:
        sys.argv = argv
    assert a.batch == 2 * a.micro_batch and bench.pick_traffic(a.micro_batch, "conv_tc_x3")[0]
