"""Multi-GPU sharding on real GPUs (skipped with fewer than 2): the frame assembled in rank 0's image by N ranks
storing over NVLink must be byte-identical to the single-GPU frame."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
    import numpy as np, torch, torch.distributed as dist
    import _pkg, scenes
    mm = _pkg.load_package()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    assets = scenes.load_assets()
    sc = scenes.make_scene(mm, "C1", assets, W=400, H=231)
    cs = mm.ComputeShader(rank, (400, 231), placement=sc["textures"]["placement"], curl=sc["textures"]["curl"],
                          lowRes=sc["textures"]["lowres"], hiRes=sc["textures"]["hires"])
    cs.setFilterMode(mm.MM_FILTER_HW)            # the default production mode
    cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    shared = mm.multigpu.SharedFrame(cs, rank, world, dist)
    dist.barrier()
    cs.dispatch(mm.MM_FULL, rank, world, 4)
    cs.synchronize()
    dist.barrier()
    # the same frame again with the shared HOST frame bound: every rank's kernel also stores its pixels into one
    # page-locked shared-memory mapping (the end-to-end path: no gather, no device->host copy afterwards)
    host = mm.multigpu.SharedHostFrame(cs, rank, world, dist)
    if rank == 0: host.array[...] = -7.0
    dist.barrier()
    cs.dispatch(mm.MM_FULL | mm.MM_ROWS_SNAKE, rank, world, 8)
    cs.synchronize()
    dist.barrier()
    host_frame = host.array.copy() if rank == 0 else None
    host.close()
    if rank == 0:
        sharded = cs.readOutput()
        assert np.array_equal(host_frame.view(np.uint32), sharded.view(np.uint32)), "host frame differs from rank 0's device image"
        cs.allocOutput()
        cs.dispatch(mm.MM_FULL)
        cs.synchronize()
        single = cs.readOutput()
        assert np.array_equal(sharded.view(np.uint32), single.view(np.uint32)), "sharded frame differs from the single-GPU frame"
        print("multigpu ok", world)
    dist.barrier()
    shared.close()
    cs.close()
    dist.destroy_process_group()
""")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_frame_is_byte_identical(tmp_path, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=root))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert f"multigpu ok {world}" in out.stdout


def test_dispatch_multi_over_two_devices_of_one_process(mm, assets):
    """several GPUs driven by ONE process -- one context per device, mm_enable_peer, the same image bound in each,
    mm_dispatch_multi -- must assemble the single-GPU frame bit for bit (without the peer call the second device's stores would fault)."""
    import torch
    import scenes
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    W, H = 320, 187
    sc = scenes.make_scene(mm, "C1", assets, W=W, H=H)
    tex = sc["textures"]
    shaders = [mm.ComputeShader(d, (W, H), placement=tex["placement"], curl=tex["curl"], lowRes=tex["lowres"], hiRes=tex["hires"]) for d in (0, 1)]
    ptr, pitch = shaders[0].allocOutput()
    shaders[1].enablePeer(shaders[0])
    shaders[1].bindOutput(ptr, pitch)
    for cs in shaders:
        cs.updateUniformBuffers(sc["cam"], None, sc["sky"], sc["sun"])
    mm.dispatchMulti(shaders, mm.MM_FULL, 8)
    for cs in shaders:
        cs.synchronize()
    sharded = shaders[0].readOutput()
    shaders[0].dispatch(mm.MM_FULL)
    shaders[0].synchronize()
    single = shaders[0].readOutput()
    for cs in shaders:
        cs.close()
    assert np.array_equal(sharded.view(np.uint32), single.view(np.uint32))
