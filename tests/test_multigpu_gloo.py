"""World-size-2 test of the multi-rank host logic on CPU (gloo): band ownership, padded gather to
rank 0 and the scatter of the gathered bands into the final frame -- the same functions bench.py
uses with NCCL.  Rendering is replaced by a deterministic pattern (no GPU here)."""
from __future__ import annotations

import os
import subprocess
import sys
import textwrap

from conftest import ROOT


def test_band_gather_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, ctypes as C, importlib
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch, torch.distributed as dist
        import bench
        capi = importlib.import_module("raytracer-in-cpp_b200").capi
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        W, H = 64, 37   # last band partial
        params = capi.make_params(W, H, band_rows=8, band_rank=rank, band_world=world)
        rows = capi.local_row_map(params)
        # "render": pixel value encodes its global row and column
        local = torch.zeros((len(rows), W, 4), dtype=torch.uint8)
        for k, r in enumerate(rows):
            local[k, :, 0] = int(r); local[k, :, 1] = torch.arange(W, dtype=torch.uint8); local[k, :, 3] = 255
        frame = bench.gather_frame(local, rows, W, H, rank, world, dist, torch, torch.device("cpu"))
        if rank == 0:
            exp = torch.zeros((H, W, 4), dtype=torch.uint8)
            exp[:, :, 0] = torch.arange(H, dtype=torch.uint8)[:, None]
            exp[:, :, 1] = torch.arange(W, dtype=torch.uint8)[None, :]
            exp[:, :, 3] = 255
            assert torch.equal(frame, exp), "gathered frame differs"
            print("GATHER_OK")
        else:
            assert frame is None
        dist.destroy_process_group()
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "GATHER_OK" in r.stdout


def test_host_frame_world2_shared_memory(tmp_path):
    """The no-collective assembly path of the per-GPU processes, on CPU: two processes open the same POSIX
    shared-memory host frame (rt_host_frame_*), each writes the rows rt_local_row_map gives it, and the frame is
    complete after the library's own process barrier -- no torch.distributed in the data path."""
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent(f"""
        import os, sys, importlib, time
        sys.path.insert(0, {ROOT!r})
        import numpy as np
        capi = importlib.import_module("raytracer-in-cpp_b200").capi
        rank, world, name = int(sys.argv[1]), 2, sys.argv[2]
        W, H = 48, 37   # last band partial
        if rank != 0:
            time.sleep(0.5)   # the creator must come first
        hf = capi.HostFrame(name, W, H, create=(rank == 0))
        for frame_no in range(3):
            hf.barrier(world)                       # previous frame consumed
            params = capi.make_params(W, H, band_rows=8, band_rank=rank, band_world=world)
            for r in capi.local_row_map(params):
                hf.array[r, :, 0] = r
                hf.array[r, :, 1] = np.arange(W)
                hf.array[r, :, 2] = frame_no
                hf.array[r, :, 3] = 255
            hf.barrier(world)                       # frame complete
            if rank == 0:
                exp = np.zeros((H, W, 4), np.uint8)
                exp[:, :, 0] = np.arange(H)[:, None]; exp[:, :, 1] = np.arange(W)[None, :]; exp[:, :, 2] = frame_no; exp[:, :, 3] = 255
                assert (hf.array == exp).all(), "assembled host frame differs"
        hf.barrier(world)
        hf.close()
        print("HOSTFRAME_OK", rank)
    """))
    name = f"/rt_test_frame_{os.getpid()}"
    procs = [subprocess.Popen([sys.executable, str(script), str(r), name], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True, cwd=ROOT) for r in range(2)]
    outs = [p.communicate(timeout=120) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o[-1000:] + e[-3000:]
        assert "HOSTFRAME_OK" in o
