"""world_size-2 gloo tests of the data-parallel host logic (no GPU): segment sharding and the flat
gradient bucket all-reduce (SUM / world), the only collective of the path."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nafae_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    b = parallel.GradBucket(parallel.trainable_grad_elems(), torch.device("cpu"), world)
    g_vis, g_bias, g_word = b.views([(512, 4096), (512,), (512, 200)])
    g_vis.fill_(float(rank + 1))
    g_bias.copy_(torch.arange(512, dtype=torch.float32) * (rank + 1))
    g_word.fill_(-2.0 * (rank + 1))
    b.allreduce_async()
    b.wait()
    mean = (1 + world) / 2.0
    ok = bool(torch.allclose(g_vis, torch.full_like(g_vis, mean)) and
              torch.allclose(g_bias, torch.arange(512, dtype=torch.float32) * mean) and
              torch.allclose(g_word, torch.full_like(g_word, -2.0 * mean)))
    # replicas agree bit-for-bit after the all-reduce
    gathered = [torch.zeros(8) for _ in range(world)]
    dist.all_gather(gathered, b.buf[:8].clone())
    ok = ok and all(torch.equal(gathered[0], g) for g in gathered)
    out[rank] = ok
    dist.destroy_process_group()


def test_grad_bucket_allreduce_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_shard_segments_partitions_everything_once():
    for n in (10000, 17, 8, 3):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_segments(n, r, world) for r in range(world)]
            covered = np.zeros(n, int)
            for b, e in spans:
                covered[b:e] += 1
            assert (covered == 1).all()
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_trainable_grad_bucket_size_matches_reference_parameter_shapes():
    # vis_ebd.fc1 (512x4096 + 512), word_ebd.fc1 (512x200 + 512), word_ebd.bn (2x512): model.py:616-642
    assert parallel.trainable_grad_elems() == 512 * 4096 + 512 + 512 * 200 + 512 + 1024
    assert abs(parallel.trainable_grad_elems() * 4 / 1e6 - 8.8) < 0.1


def _sweep_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_eval_golden import make_case, to_reference_inputs
    from nafae_b200 import evaluate
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    # 90 single-frame "segments": every rank records the detections of its own contiguous shard
    z = make_case(seed=21, n_imgs=90, n_cls=9)
    recs, dets, class_list = to_reference_inputs(z)
    begin, end = parallel.shard_segments(90, rank, world)
    mine = [k for k, img in enumerate(dets[0]) if begin <= img < end]
    local = [[lst[k] for k in mine] for lst in dets]
    merged = parallel.gather_dets(local)
    whole = evaluate.box_accuracy_details(recs, dets, class_list)
    got = evaluate.box_accuracy_details(recs, merged, class_list)
    out[rank] = bool(merged[0] == dets[0] and merged[1] == dets[1] and
                     np.array_equal(got["class_match_count"], whole["class_match_count"]) and
                     got["macro"] == whole["macro"] and
                     evaluate.phrase_accuracy_details(recs, merged, class_list)["macro"] ==
                     evaluate.phrase_accuracy_details(recs, dets, class_list)["macro"])
    dist.destroy_process_group()


def test_sharded_inference_sweep_merges_to_the_single_process_result_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sweep_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
