"""Host side of the multi-GPU path.
CPU (gloo, world_size 2): strip partition and the IPC-handle exchange.  GPU (only on boxes with >= 2 GPUs): the
cross-process check of tests/multi_gpu_check.py under torchrun."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_partition_rows_covers_the_frame_with_whole_tiles():
    from raytracingdenoiser_b200 import strips
    for h in (16, 141, 720, 1080, 2160, 4320):
        for n in (1, 2, 3, 4, 8):
            tiles = (h + 15) // 16
            per_rank = (tiles + n - 1) // n
            if per_rank * (n - 1) >= tiles:          # the last rank(s) would own nothing
                with pytest.raises(ValueError):
                    strips.partition_rows(h, n)
                continue
            s, parts = strips.partition_rows(h, n)
            assert s % 16 == 0 and len(parts) == n
            assert parts[0][0] == 0 and parts[-1][1] == h
            for r, (y0, y1) in enumerate(parts):
                assert y0 == r * s and 0 < y1 - y0 <= s
                if r + 1 < n:
                    assert y1 == parts[r + 1][0] and y1 - y0 == s
    assert strips.partition_rows(2160, 8) == (272, [(r * 272, min(2160, (r + 1) * 272)) for r in range(8)])


def test_partition_rows_weighted_balances_cost_and_respects_minimum():
    from raytracingdenoiser_b200 import strips
    tiles = 135                                  # 2160 rows
    cost = [1.0] * 18 + [60.0] * (tiles - 18)    # 13 % sky on top
    for n in (2, 4, 8):
        cap, parts = strips.partition_rows_weighted(2160, n, cost, min_rows=96)
        assert parts[0][0] == 0 and parts[-1][1] == 2160 and cap % 16 == 0
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(y0 % 16 == 0 and y1 - y0 >= 96 and y1 - y0 <= cap for y0, y1 in parts)
        shares = [sum(cost[y0 // 16:(y1 + 15) // 16]) for y0, y1 in parts]
        assert max(shares) <= sum(cost) / n + 2 * 60.0   # within a tile row or two of the ideal share
        assert parts[0][1] - parts[0][0] > parts[-1][1] - parts[-1][0]   # the sky strip is taller
    with pytest.raises(ValueError):
        strips.partition_rows_weighted(64, 8, [1.0] * 4)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from raytracingdenoiser_b200 import strips
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([(rank * 37 + i) & 255 for i in range(64)])
    got = strips.exchange_ipc_handles(mine)
    q.put((rank, got))
    dist.barrier()
    dist.destroy_process_group()


def test_ipc_handle_exchange_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [bytes([(r * 37 + i) & 255 for i in range(64)]) for r in range(2)]
    assert res[0] == expect and res[1] == expect


@pytest.mark.gpu
@pytest.mark.parametrize("denoiser", ["REBLUR_DIFFUSE_SPECULAR", "RELAX_DIFFUSE_SPECULAR", "SIGMA_SHADOW"])
def test_cross_process_strips_bit_identical(denoiser):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(HERE, "multi_gpu_check.py"), denoiser, "640", "368", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_measured_rebalancing_moves_rows_away_from_the_slow_rank():
    """strips.rebalance_tile_row_cost: per-rank kernel times of a trial partition correct the per-tile-row costs; cutting again moves
    the boundaries towards equal time (bench.py does this during set-up at N > 1)."""
    from raytracingdenoiser_b200 import strips
    h, world = 2160, 4
    cost = [1.0] * ((h + 15) // 16)
    cap, part = strips.partition_rows_weighted(h, world, cost, min_rows=96)
    # rows of rank 0 cost twice as much per row as the model thinks
    ms = [2.0 * (y1 - y0) if r == 0 else 1.0 * (y1 - y0) for r, (y0, y1) in enumerate(part)]
    cost2 = strips.rebalance_tile_row_cost(cost, part, ms)
    cap2, part2 = strips.partition_rows_weighted(h, world, cost2, min_rows=96)
    assert part2[0][1] - part2[0][0] < part[0][1] - part[0][0]
    assert part2[0][0] == 0 and part2[-1][1] == h and all(a[1] == b[0] for a, b in zip(part2, part2[1:]))
    assert all(y0 % 16 == 0 for y0, _ in part2) and cap2 >= max(y1 - y0 for y0, y1 in part2)

    def time_of(p):  # what the strips would take under the true cost density
        return [sum((2.0 if y < part[0][1] else 1.0) for y in range(y0, y1)) for y0, y1 in p]
    assert max(time_of(part2)) < 0.8 * max(time_of(part))
    # balanced measurements leave the costs proportional to what they were
    same = strips.rebalance_tile_row_cost(cost, part, [float(y1 - y0) for y0, y1 in part])
    assert max(abs(a - b) for a, b in zip(same, cost)) < 1e-6
