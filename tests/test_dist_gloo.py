"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the document sharding plan, the packed
64-bit keys and the all-reduce-max that merges per-shard longest-suffix matches.
Per-shard lookups come from the CPU oracle here; the GPU lookups are covered by tests -m gpu."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import samd_oracle as O
from samd_b200 import dist as D
from samd_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _corpus():
    docs = [d.tolist() for d in synth.make_corpus(6000, 300, 51, doc_len=(16, 64), singletons=True)]
    return docs, synth.EOS


def _queries(docs, n, length, seed):
    q = synth.corpus_queries([np.array(d) for d in docs], n, length, 300, seed)
    q[q == synth.EOS] = 7                       # EOS-free: equality with the single global automaton is exact
    return q


def _worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    docs, eos = _corpus()
    plan = D.shard_documents([len(d) for d in docs], [d[-1] == eos for d in docs], world)
    lo, hi, offset = plan[rank]
    shard = O.build_static(docs[lo:hi], eos)
    q = _queries(docs, 48, 24, 52)
    keys = np.zeros(len(q), dtype=np.int64)
    for i, row in enumerate(q):
        shard.reset_cursor()
        shard.advance(row[:-1])
        state, length = shard.peek(int(row[-1]))
        keys[i] = D.pack_key(np.array([length]), np.array([offset + shard.first_end[state]]))[0]
    t = torch.from_numpy(keys)
    D.reduce_keys(t)
    out_q.put((rank, t.numpy().copy(), plan))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_lookup_allreduce_max_matches_global_automaton():
    world = 2
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out_q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out_q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    results.sort(key=lambda x: x[0])
    assert np.array_equal(results[0][1], results[1][1])                  # every rank holds the reduced keys
    docs, eos = _corpus()
    whole = O.build_static(docs, eos)
    q = _queries(docs, 48, 24, 52)
    lengths, ends = D.unpack_key(results[0][1])
    text = D.flatten_corpus(docs, eos)
    assert len(text) - 1 == whole.n
    n_hit = 0
    for i, row in enumerate(q):
        whole.reset_cursor()
        whole.advance(row[:-1])
        state, length = whole.peek(int(row[-1]))
        assert lengths[i] == length
        if length:
            n_hit += 1
            assert ends[i] == whole.first_end[state]
            e = int(ends[i])
            assert text[e + 1:e + 16].tolist() == whole.text[e + 1:e + 16]     # the draft both would read
    assert n_hit > 10


def test_shard_plan_covers_documents_and_offsets():
    rng = np.random.default_rng(3)
    lens = rng.integers(1, 50, size=200)
    ends = rng.random(200) < 0.3
    for world in (1, 2, 3, 4, 8):
        plan = D.shard_documents(lens, ends, world)
        assert plan[0][0] == 0 and plan[-1][1] == 200
        tok = lens + (~ends)
        for g in range(world):
            lo, hi, off = plan[g]
            assert off == int(tok[:lo].sum())
            if g:
                assert plan[g - 1][1] == lo
        sizes = [int(tok[lo:hi].sum()) for lo, hi, _ in plan]
        assert max(sizes) - min(sizes) <= 2 * int(tok.max())


def test_key_packing_orders_by_length_then_earliest_position():
    k = D.pack_key(np.array([3, 3, 4, 0]), np.array([100, 50, 1000, 5]))
    assert k[3] == 0 and k[2] > k[1] > k[0] > 0
    length, end = D.unpack_key(k)
    assert length.tolist() == [3, 3, 4, 0] and end.tolist() == [100, 50, 1000, 0]
    kt = D.pack_key(torch.tensor([3, 0]), torch.tensor([7, 9]))
    assert D.unpack_key(kt)[1].tolist() == [7, 0]
