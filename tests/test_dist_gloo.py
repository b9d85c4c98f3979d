"""CPU, world_size 2 over gloo: the N>1 host logic (sharding, ragged gather, bucketed gradient averaging)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # dist.py has no dependency on the CUDA library: load it standalone so this test needs no build
    spec = importlib.util.spec_from_file_location("agb_dist", os.path.join(root, "autognothi_b200", "dist.py"))
    agd = importlib.util.module_from_spec(spec)
    sys.modules["agb_dist"] = agd
    spec.loader.exec_module(agd)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. sharding: disjoint, contiguous, covering, balanced
        n = 37
        lo, hi = agd.shard_range(n, rank, world)
        spans = [agd.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n and all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
        # 2. ragged gather keeps global row order b*S+s
        S, C = 4, 3
        counts = [(b - a) * S for a, b in spans]
        local = torch.arange(lo * S, hi * S, dtype=torch.float32)[:, None].repeat(1, C)
        full = agd.gather_rows(local, counts)
        assert full.shape == (n * S, C) and torch.equal(full[:, 0], torch.arange(n * S, dtype=torch.float32))
        # 3. bucketed gradient averaging == mean over ranks, for several bucket sizes (incl. one-param buckets)
        torch.manual_seed(0)
        shapes = [(5, 7), (11,), (3, 2, 2), (1,), (64, 9)]
        for bucket_mb in (1e-5, 1e-3, 64.0):
            params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes]
            for i, p in enumerate(params):
                p.grad = torch.full(p.shape, float(rank + 1) * (i + 1))
            params[3].grad = None if rank == 0 else params[3].grad  # a rank without a gradient contributes zeros
            red = agd.GradAllReducer(params, bucket_mb=bucket_mb)
            assert sum(len(b) for b in red.buckets) == len(params)
            red.allreduce()
            for i, p in enumerate(params):
                want = (i + 1) * sum(r + 1 for r in range(world)) / world
                if i == 3:
                    want = (i + 1) * sum(r + 1 for r in range(1, world)) / world
                assert torch.allclose(p.grad, torch.full(p.shape, want)), (bucket_mb, i, p.grad.flatten()[:3], want)
        # 4. overlapped reducer: gradients pushed stage by stage (as the hand-written adjoint does), several bucket sizes,
        #    both wire formats; step 1 records the layout, steps 2-3 reuse it and launch buckets as they fill
        names = [f"p{i}" for i in range(len(shapes))]
        for bucket_mb, wire in ((1e-5, torch.float32), (2e-4, torch.float32), (32.0, torch.float32), (2e-4, torch.bfloat16)):
            red = agd.OverlappedGradReducer(bucket_mb=bucket_mb, wire_dtype=wire)
            for step in range(3):
                grads = {}
                for stage in ((0, 1), (2,), (3, 4)):
                    for i in stage:
                        grads[names[i]] = torch.full(shapes[i], float(rank + 1) * (i + 1) + step)
                    red.push(grads, [names[i] for i in stage])
                out = red.finish(grads)
                for i, nme in enumerate(names):
                    want = (i + 1) * sum(r + 1 for r in range(world)) / world + step
                    assert out[nme].shape == torch.Size(shapes[i]) and out[nme].dtype == torch.float32
                    assert torch.allclose(out[nme], torch.full(shapes[i], want), rtol=1e-2 if wire == torch.bfloat16 else 1e-6), \
                        (bucket_mb, wire, step, nme, out[nme].flatten()[:3], want)
            assert len(red.flat) >= (3 if bucket_mb < 1e-4 else 1)
        # 5. parameter broadcast + replica check
        torch.manual_seed(100 + rank)
        ps = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
        try:
            agd.assert_replicas_identical(ps)
            raise AssertionError("diverged replicas not detected")
        except RuntimeError:
            pass
        agd.broadcast_parameters(ps, src=0)
        agd.assert_replicas_identical(ps)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, f"FAIL {type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_shard_range_single_process():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("agb_dist1", os.path.join(root, "autognothi_b200", "dist.py"))
    agd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(agd)
    for n in (0, 1, 7, 8, 1000):
        for w in (1, 2, 3, 8):
            spans = [agd.shard_range(n, r, w) for r in range(w)]
            assert sum(b - a for a, b in spans) == n
            assert all(a <= b for a, b in spans)
