"""Host logic of the utterance-sharded multi-GPU path, incl. a real world_size-2 gloo run on CPU."""
import os
import random

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from autoformer_b200 import sharding


def _lengths(n, seed=0):
    rng = random.Random(seed)
    return [rng.choice(range(128, 1025, 32)) for _ in range(n)]


def test_buckets_are_exact_length_and_cover_everything():
    lens = _lengths(5000)
    b = sharding.bucket_by_length(lens)
    assert all(t % 32 == 0 for t in b) and sum(len(v) for v in b.values()) == 5000
    for t, ids in b.items():
        assert all(lens[i] == t for i in ids)
    batches = sharding.make_batches(b, 512)
    assert all(1 <= len(ids) <= 512 for _, ids in batches)
    assert sorted(i for _, ids in batches for i in ids) == list(range(5000))
    assert sharding.make_batches({}, 512) == []                         # empty input


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_assignment_is_balanced_and_deterministic(world):
    lens = _lengths(65536 // 8, seed=1)
    plan_a = sharding.plan(lens, world)
    plan_b = sharding.plan(lens, world)
    assert plan_a == plan_b
    seen = sorted(i for r in plan_a for _, ids in r for i in ids)
    assert seen == list(range(len(lens)))                               # every utterance exactly once
    load = [sum(t * len(ids) for t, ids in r) for r in plan_a]
    assert max(load) - min(load) <= 1024 * 512                          # within one largest batch
    assert max(load) <= sum(load) / world * 1.15


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = _lengths(300, seed=2)
    mine = sharding.plan(lens, world, max_batch=16)[rank]
    ids = [i for _, b in mine for i in b]
    frames = float(sum(t * len(b) for t, b in mine))
    # stand-in for the conversion: output row = [id, length]
    local = torch.tensor([[float(i), float(lens[i])] for i in ids]).reshape(-1, 2)
    recs = sharding.gather_records([frames, float(len(ids)), float(local.sum())])
    full = sharding.gather_outputs(local, len(lens), ids, dst=0)
    if rank == 0:
        q.put((recs.tolist(), full.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + random.randint(0, 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    recs, full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lens = _lengths(300, seed=2)
    assert len(recs) == 2 and sum(r[1] for r in recs) == 300
    assert sum(r[0] for r in recs) == float(sum(lens))
    assert [int(v[0]) for v in full] == list(range(300))                # every utterance landed at its own index
    assert [int(v[1]) for v in full] == lens


def test_cost_model_balances_tail_batches():
    """LPT by the measured AutoVC cost model (LSTM time per frame is independent of the batch size up to 512) must
    balance the modelled time better than LPT by frame count, and still cover every utterance exactly once."""
    import random
    from autoformer_b200 import sharding
    rng = random.Random(1234)
    lengths = [rng.choice(range(128, 1025, 32)) for _ in range(8192)]
    spread = {}
    for cost in (sharding.frames_cost, sharding.autovc_cost):
        plan = sharding.plan(lengths, 8, 512, cost=cost)
        seen = sorted(i for rank in plan for _, ids in rank for i in ids)
        assert seen == list(range(len(lengths)))
        loads = [sum(sharding.autovc_cost(t, len(ids)) for t, ids in rank) for rank in plan]
        spread[cost.__name__] = max(loads) / (sum(loads) / len(loads))
    assert spread["autovc_cost"] <= spread["frames_cost"] + 1e-9
    assert spread["autovc_cost"] < 1.05


def test_cost_model_knows_that_tail_batches_run_in_pairs():
    """pipeline.convert_batches runs batches of 65 .. 256 utterances two at a time: the LPT cost of such a batch carries
    0.6 of its frame time; full batches, weight-stationary-sized ones and pair_below = 0 keep the whole of it."""
    from autoformer_b200 import sharding
    full = sharding.autovc_cost(1024, 512, "fp16x2")
    tail = sharding.autovc_cost(1024, 212, "fp16x2")
    tail_alone = sharding.autovc_cost(1024, 212, "fp16x2", pair_below=0)
    per_frame, per_utt_frame = sharding._AUTOVC_COST["fp16x2"]
    assert abs(tail_alone - 1024 * (per_frame + per_utt_frame * 212)) < 1e-6
    assert abs(tail - 1024 * (sharding.PAIRED_LSTM_SHARE * per_frame + per_utt_frame * 212)) < 1e-6
    assert tail < tail_alone < full
    assert sharding.autovc_cost(1024, 64, "fp16x2") == sharding.autovc_cost(1024, 64, "fp16x2", pair_below=0)
    assert sharding.autovc_cost(1024, 300, "fp16x2") == sharding.autovc_cost(1024, 300, "fp16x2", pair_below=0)
    # the plan stays a partition of the utterances and deterministic
    lengths = [128 + 32 * (i % 29) for i in range(5000)]
    a = sharding.plan(lengths, 8, 512, cost=sharding.autovc_cost_for("fp16x2"))
    b = sharding.plan(lengths, 8, 512, cost=sharding.autovc_cost_for("fp16x2"))
    assert a == b
    ids = sorted(i for r in a for _, batch in r for i in batch)
    assert ids == list(range(5000))
    loads = [sum(sharding.autovc_cost(t, len(batch), "fp16x2") for t, batch in r) for r in a]
    assert max(loads) / (sum(loads) / len(loads)) < 1.1
