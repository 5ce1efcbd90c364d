"""Host-side logic of the multi-GPU path on CPU: the row-cyclic partition and the handle exchange over a
world_size-2 gloo process group (the kernels' half of the contract is tested on the GPU by
test_row_partition_union_equals_full_frame)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest


def test_partition_is_an_exact_cover(mm):
    for H in (1, 7, 117, 1080, 2160):
        for world in (1, 2, 3, 8):
            for block in (1, 2, 4, 16):
                assert mm.multigpu.partition_is_exact_cover(H, world, block), (H, world, block)
                assert mm.multigpu.partition_is_exact_cover(H, world, block, snake=True), (H, world, block, "snake")


def test_partition_matches_oracle_dispatch(mm, oracle, assets):
    """owned_rows() == the rows the oracle's row partition writes (same rule as mm_dispatch)."""
    import scenes
    W, H = 24, 37
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    S = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"])
    for world, block in ((2, 1), (3, 2), (4, 4)):
        for r in range(world):
            out, _ = S.march(W, H, row_begin=r, row_stride=world, row_block=block, counters=False, out=np.full((H, W, 4), -7, np.float32))
            rows = np.nonzero((out != -7).any(axis=(1, 2)))[0]
            assert np.array_equal(rows, mm.multigpu.owned_rows(H, r, world, block))


def test_row_cyclic_balances_load_better_than_bands(mm, oracle, assets):
    """rows below the horizon are free, so contiguous bands are badly unbalanced; row-cyclic is not."""
    import scenes
    W, H = 160, 90
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    _, cnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(W, H)
    row_cost = (cnt[..., 1] + 2 * cnt[..., 2]).sum(axis=1).astype(np.float64)
    n = 8
    cyc = [row_cost[mm.multigpu.owned_rows(H, r, n, 2)].sum() for r in range(n)]
    band = [c.sum() for c in np.array_split(row_cost, n)]
    eff = lambda parts: np.mean(parts) / np.max(parts)
    assert eff(cyc) > 0.85 and eff(cyc) > eff(band) + 0.2


def test_snake_order_removes_the_rank_bias_of_coarse_row_blocks(mm, oracle, assets):
    """With 8-row blocks (the kernels' tile height) the plain cyclic order gives the last rank of every round the rows nearest
    the horizon; running odd rounds in reverse rank order (MM_ROWS_SNAKE) cancels that gradient."""
    import scenes
    W, H = 240, 540                     # 4K's row count / 4: 8 ranks x 8 rows = 64-row rounds, as on the real frame
    sc = scenes.make_scene(mm, "C3", assets, W=W, H=H)
    _, cnt = oracle.Scene(sc["textures"], sc["cam"], sc["sun"], sc["sky"]).march(W, H)
    row_cost = (cnt[..., 1] + 2 * cnt[..., 2]).sum(axis=1).astype(np.float64)
    eff = lambda parts: np.mean(parts) / np.max(parts)
    plain = [row_cost[mm.multigpu.owned_rows(H, r, 8, 8)].sum() for r in range(8)]
    snake = [row_cost[mm.multigpu.owned_rows(H, r, 8, 8, snake=True)].sum() for r in range(8)]
    assert eff(snake) > eff(plain) and eff(snake) > 0.93, (eff(plain), eff(snake))


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    import _pkg
    mm = _pkg.load_package()
    dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
    rank = dist.get_rank()
    handle = bytes(range(64)) if rank == 0 else None
    got = mm.multigpu.exchange_handle(handle, rank, 2, dist)
    assert got == bytes(range(64)), got
    for snake in (False, True):
        rows = mm.multigpu.owned_rows(37, rank, 2, 2, snake)
        gathered = [None, None]
        dist.all_gather_object(gathered, rows.tolist())
        assert sorted(gathered[0] + gathered[1]) == list(range(37))
    # the shared HOST frame of a sharded frame: one shared-memory mapping opened by both processes (the CUDA page-locking and
    # the kernel-side stores are the GPU half, tests/test_multigpu_gpu.py); here every rank writes its own rows from the CPU
    class FakePass:
        width, height = 40, 37
        def hostRegister(self, a): pass
        def hostUnregister(self, a): pass
        def bindHostMirror(self, a): pass
    host = mm.multigpu.SharedHostFrame(FakePass(), rank, 2, dist)
    assert host.array.shape == (37, 40, 4)
    assert rank != 0 or not os.path.exists(host.path)      # rank 0 unlinks the name once every rank has the file open
    host.array[mm.multigpu.owned_rows(37, rank, 2, 8)] = float(rank + 1)
    dist.barrier()
    import numpy as np
    want = np.zeros(37)
    for r in range(2):
        want[mm.multigpu.owned_rows(37, r, 2, 8)] = r + 1
    assert (host.array[:, 0, 0] == want).all() and (host.array == host.array[:, :1, :1]).all()
    dist.barrier()
    host.close()
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def test_handle_exchange_over_gloo_world_size_2(tmp_path):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=root))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out
        assert "ok" in out


def _barrier_worker(path, rank, world, rounds, q):
    import mmap
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import _pkg
    mg = _pkg.load_package().multigpu
    f = open(path, "r+b")
    m = mmap.mmap(f.fileno(), 4096)
    b = mg.HostBarrier(m, rank, world, offset=1024)
    log = np.frombuffer(m, dtype=np.int64, count=world, offset=0)
    ok = True
    for r in range(1, rounds + 1):
        log[rank] = r                      # "my part of frame r is in the shared frame"
        b.wait()
        ok = ok and bool((log >= r).all())  # after the barrier every rank's part of frame r is visible
        b.wait()                           # nobody starts frame r + 1 before everybody has checked frame r
    b.release()
    del log
    q.put((rank, ok))


def test_host_barrier_orders_the_ranks(tmp_path):
    """multigpu.HostBarrier: the shared-memory rendezvous the sharded end-to-end path uses once per frame instead of a collective."""
    import multiprocessing as mp
    path = str(tmp_path / "barrier.bin")
    with open(path, "wb") as f:
        f.truncate(4096)
    world, rounds = 3, 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_barrier_worker, args=(path, r, world, rounds, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=30)
    assert res == [(r, True) for r in range(world)]


def test_host_barrier_fails_instead_of_hanging_when_a_rank_is_missing(mm):
    """A peer process that died must turn into an error on the surviving ranks, not into a spin until the job's time limit."""
    import time
    buf = bytearray(4096)
    b = mm.multigpu.HostBarrier(buf, 0, 2, offset=0, timeout_s=0.2)
    t0 = time.monotonic()
    with pytest.raises(RuntimeError, match=r"ranks \[1\] did not reach epoch 1"):
        b.wait()
    assert time.monotonic() - t0 < 5.0
    b.release()
